#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-q}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -x -k "golden or edge or empty" > gpurun_out/${TAG}_pytest_subset.log 2>&1; echo "subset rc=$?"; grep -E "bf16x3\]|passed|failed|Error" gpurun_out/${TAG}_pytest_subset.log | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
