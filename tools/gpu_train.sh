#!/bin/bash
# Training path on the GPU box: sparse-encoder training step, render backward, generator-level training step, training-view timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-train}
timeout 900 python -m pytest tests/test_sparse_encoder.py tests/test_backward_gpu.py tests/test_training_gpu.py -m gpu -q -s > gpurun_out/${TAG}_pytest_training.log 2>&1; echo "pytest rc=$?"
grep -E "worst|passed|failed|Error|error|rel L2 [0-9.e-]+ *$|training step|sparse encoder train" gpurun_out/${TAG}_pytest_training.log | tail -30
timeout 300 python tools/bench_backward.py > gpurun_out/${TAG}_bench_backward.json 2> gpurun_out/${TAG}_bench_backward.err; echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench_backward.json
BWD_STEPS=1 BWD_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python tools/bench_backward.py > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv 12
