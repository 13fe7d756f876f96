#!/bin/bash
# One gpurun call that refreshes the round-2 evidence under gpurun_out/: GPU tests (full log), the default bench line (with the reference
# CPU arm), the importance / reference / GPU-eager lines, the ncu launch list of the bench command and one --set full capture of the
# first launch of every point-stage kernel (524 288 points) + the two cull kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2ev}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --importance 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_importance64.json 2>> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py imp64 < gpurun_out/${TAG}_bench_importance64.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gpu-eager-baseline > gpurun_out/${TAG}_bench_gpu_eager.json 2>> gpurun_out/${TAG}_bench.err; echo "eager rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench_gpu_eager.json').read().splitlines() if l.startswith('{')][-1])
print('gpu_eager', d.get('gpu_eager_baseline'))
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv 14
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_front_fused|^k_xformer_bf16|k_decoder_pp|k_cull_search|k_cull_candidates' -s 0 -c 5 -f \
  -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-200
