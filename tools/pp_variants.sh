#!/bin/bash
# timing experiments on the ping-pong decoder's MMA stream (results of variants != 0 are numerically wrong on purpose)
cd "$(dirname "$0")/.."
for v in 0 1 2 3; do
  echo "== variant $v"
  SHERF_PP_VARIANT=$v timeout 300 python tools/trace_fused.py bf16x3 2>&1 | tail -1
done
