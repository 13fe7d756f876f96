#!/bin/bash
# one-shot GPU check of the ping-pong bf16x3 decoder: parity on the golden fixtures (both packing variants), then benches
cd "$(dirname "$0")/.."
for v in 0 1; do
  echo "== variant $v"
  SHERF_PP_VARIANT=$v timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -s -k "golden and bf16x3" 2>&1 | grep -E "^\[|passed|failed|Error|error" | head -12
done
echo "== bench bf16x3"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16x3 2>&1 | tail -1 | python tools/bench_brief.py
echo "== cap sweep (tf32x3)"
for c in 131072 262144 1048576; do
  SHERF_CHUNK_CAP=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_brief.py $c
done
