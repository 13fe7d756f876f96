#!/bin/bash
# ncu --set full with source for the new point-stage kernels (one 524 288-point launch each) + full GPU suite with the async first chunk
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_front_fused|k_xformer_bf16' -s 2 -c 2 -f \
  -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q -s -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
