#!/bin/bash
# ncu --set full with source for the secondary kernels (grid build, cull, fusion, transformer), one launch each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-p2}
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_build_grids|k_cull|k_fusion_fused|k_xformer_fused|k_to_channels_last' -s 12 -c 12 -f \
  -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-200
