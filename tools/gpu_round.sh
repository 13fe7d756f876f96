#!/bin/bash
# full GPU regression + bench sweep (chunk cap x precision)
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for c in 131072 524288 1048576; do
  SHERF_CHUNK_CAP=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16x3 2>&1 | tail -1 | python tools/bench_brief.py bf16x3 cap=$c
done
SHERF_CHUNK_CAP=1048576 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision tf32x3 2>&1 | tail -1 | python tools/bench_brief.py tf32x3 cap=1048576
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision tf32x3 2>&1 | tail -1 | python tools/bench_brief.py tf32x3 cap=default
