#!/bin/bash
# quick GPU regression: all -m gpu tests, default bench line, configs[4]-shape (importance) bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-chk}
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "^\[|passed|failed|Error|error|assert" | tail -45 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log | tail -22
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json
SHERF_NO_PROLOGUE_OVERLAP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_brief.py c2-no-overlap
timeout 300 python bench.py --importance 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_imp.json 2>> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py c5 < gpurun_out/${TAG}_bench_imp.json; tail -3 gpurun_out/${TAG}_bench.err
