#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-b}
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/${TAG}_bench.json | python tools/bench_brief.py c2; done
SHERF_NO_PROLOGUE_OVERLAP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_brief.py c2-no-side-stream
SHERF_NO_PACK_REUSE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_brief.py c2-no-pack-reuse
