#!/bin/bash
# ncu --set full of one 128 x 128 weight-gradient launch and the dX launch that follows it (first backward chunk, 262 144 points)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-bwdncu}
BWD_STEPS=1 BWD_WARMUP=0 timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_umma_grad_w|k_umma_linear' -s ${2:-22} -c ${3:-2} -f \
  -o gpurun_out/${TAG}_full python tools/bench_backward.py > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-200
