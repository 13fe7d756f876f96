#!/bin/bash
# One gpurun call that refreshes every piece of evidence under gpurun_out/: GPU tests, the default bench line,
# the reference arm, the ncu launch list of the same command and one --set full capture of the four point-stage kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-ev}
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "^\[|passed|failed|Error|error|assert" | tail -60 > gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --importance 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_importance64.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_decoder_pp|k_xformer_fused|k_fusion_fused|k_point_gather4|k_cull_search|k_cull_candidates' -s 12 -c 6 -f \
  -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-200
ls -la gpurun_out | tail -12
