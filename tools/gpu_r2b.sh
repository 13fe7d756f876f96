#!/bin/bash
# round-2 kernel bring-up: parity subset first (fail fast, bounded), then A/B bench lines (new front / transformer kernels vs legacy)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -x -k "golden or edge or empty" > gpurun_out/${TAG}_pytest_subset.log 2>&1; echo "subset rc=$?"; grep -E "^\[|passed|failed|Error" gpurun_out/${TAG}_pytest_subset.log | tail -30
for mode in new legacy_front legacy_both; do
  export SHERF_LEGACY_FRONT= SHERF_LEGACY_XFORMER=
  unset SHERF_LEGACY_FRONT SHERF_LEGACY_XFORMER
  [ $mode = legacy_front ] && export SHERF_LEGACY_FRONT=1
  [ $mode = legacy_both ] && export SHERF_LEGACY_FRONT=1 SHERF_LEGACY_XFORMER=1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${mode}.json 2> gpurun_out/${TAG}_bench_${mode}.err
  echo "== $mode rc=$?"; python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench_${mode}.json; tail -2 gpurun_out/${TAG}_bench_${mode}.err
done
