#!/bin/bash
# backward-pass check on the GPU box: gradient parity tests, a parity subset of the forward (the gather kernel template changed), timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-bwd}
timeout 900 python -m pytest tests/test_backward_gpu.py -m gpu -q -s -x > gpurun_out/${TAG}_pytest_backward.log 2>&1; echo "backward rc=$?"
grep -E "rel L2|worst|passed|failed|Error|error|loss oracle" gpurun_out/${TAG}_pytest_backward.log | tail -${2:-60}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or edge or empty" > gpurun_out/${TAG}_pytest_subset.log 2>&1; echo "subset rc=$?"; tail -3 gpurun_out/${TAG}_pytest_subset.log
if [ -f tools/bench_backward.py ]; then timeout 600 python tools/bench_backward.py > gpurun_out/${TAG}_bench_backward.json 2> gpurun_out/${TAG}_bench_backward.err; cat gpurun_out/${TAG}_bench_backward.json; tail -3 gpurun_out/${TAG}_bench_backward.err; fi
