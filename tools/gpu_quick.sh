#!/bin/bash
# fastest GPU regression: the parity tests that exist since r1_p (no full-size oracle subsets) + the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 900 python -m pytest tests -m gpu -q -x -k "not ray_subset" 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
