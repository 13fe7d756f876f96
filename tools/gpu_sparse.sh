#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sparse_encoder.py -m gpu -q -x -s 2>&1 | grep -E "^\[|passed|failed|Error|error|assert|Traceback" | head -30
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sparse_encoder.py -m gpu -q -x -k "32-40" > gpurun_out/sparse_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sparse_memcheck.log | tail -3
