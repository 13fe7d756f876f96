#!/bin/bash
# round-2 first GPU pass: whole GPU suite (verbose prints into a log), then the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2a}
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/${TAG}_pytest_gpu.log; tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
