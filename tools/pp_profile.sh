#!/bin/bash
# cycle-counter trace + one ncu --set full capture of the ping-pong decoder (library must be built with SHERF_FUSED_TRACE=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/trace_fused.py bf16x3 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decoder_pp -s 2 -c 1 -f -o gpurun_out/prof_pp \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision bf16x3 > gpurun_out/ncu_pp.log 2>&1
tail -2 gpurun_out/ncu_pp.log | cut -c1-300
