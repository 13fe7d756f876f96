#!/bin/bash
# quick GPU check of the ping-pong decoder: golden parity, cycle trace (if built with SHERF_FUSED_TRACE=1), bench
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -s -k "golden and bf16x3" 2>&1 | grep -E "^\[|passed|failed|rror" | head -8
timeout 300 python tools/trace_fused.py bf16x3 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16x3 2>&1 | tail -1 | python tools/bench_brief.py bf16x3
