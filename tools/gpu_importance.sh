#!/bin/bash
# GPU check of the importance (fine) pass: parity tests, a memcheck run of the small cases, the configs[4]-shape bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_importance_gpu.py -m gpu -q -x -s 2>&1 | grep -E "^\[|passed|failed|Error|error|assert" | head -40
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_importance_gpu.py -m gpu -q -x -k "edge or empty or sample_importance" > gpurun_out/imp_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/imp_memcheck.log | tail -3
timeout 300 python bench.py --importance 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/imp_bench.json 2> gpurun_out/imp_bench.err
python tools/bench_brief.py c5 < gpurun_out/imp_bench.json; tail -3 gpurun_out/imp_bench.err
