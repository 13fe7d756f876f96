#!/bin/bash
# Backward pass on the GPU box: gradient parity tests, training-step timing (tensor-core and fp32 SIMT arithmetic), ncu launch list of one step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-bwd}
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -s > gpurun_out/${TAG}_pytest_backward.log 2>&1; echo "pytest rc=$?"
grep -E "worst|passed|failed|Error|error" gpurun_out/${TAG}_pytest_backward.log | tail -12
timeout 300 python tools/bench_backward.py > gpurun_out/${TAG}_bench_backward.json 2> gpurun_out/${TAG}_bench_backward.err; echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench_backward.json
if [ "${BWD_SKIP_SIMT:-1}" != "1" ]; then
SHERF_BWD_SIMT=1 timeout 300 python tools/bench_backward.py > gpurun_out/${TAG}_bench_backward_simt.json 2>> gpurun_out/${TAG}_bench_backward.err; tail -1 gpurun_out/${TAG}_bench_backward_simt.json
fi
BWD_STEPS=1 BWD_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python tools/bench_backward.py > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv 30
