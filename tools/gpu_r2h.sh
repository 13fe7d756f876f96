#!/bin/bash
# 2-GPU validation: NCCL parity test, bench at N=2 (views) and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -s > gpurun_out/${TAG}_pytest_dist.log 2>&1; echo "dist rc=$?"; grep -E "nccl|passed|failed|DIST" gpurun_out/${TAG}_pytest_dist.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
echo "n2 rc=$?"; python tools/bench_brief.py n2 < gpurun_out/${TAG}_bench_n2.json; tail -3 gpurun_out/${TAG}_bench_n2.err

python -c "
import json,sys
d=json.load(open('gpurun_out/${TAG}_bench_n2.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
print('per_rank',json.dumps(d.get('per_rank_ms'))[:600])
print('strong',json.dumps(d.get('strong_scaling_one_view')))
"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "n1 rc=$?"; python tools/bench_brief.py n1 < gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/${TAG}_bench_n1.json'))
print('cpu_baseline',d.get('cpu_baseline')); print('psnr',d.get('psnr_vs_reference')); print('roofline',json.dumps(d['roofline'])[:900])
"
