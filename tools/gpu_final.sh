#!/bin/bash
# last check of a round: every GPU test, smoke(), the default bench line, the sparse-encoder timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-final}
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
timeout 300 python tools/time_sparse_encoder.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_sparse_encoder.json
