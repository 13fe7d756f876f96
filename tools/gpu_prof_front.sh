#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-pf}
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_front_fused|k_xformer_bf16' -s 0 -c 2 -f \
  -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -1 gpurun_out/${TAG}_full.log | cut -c1-200
