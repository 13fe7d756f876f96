#!/bin/bash
# 8-GPU scaling validation: N = 8 (and 4) bench lines with the per-rank split, plus the NCCL parity test on 8 ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2j}
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
echo "n$n rc=$?"
python -c "
import json
txt=open('gpurun_out/${TAG}_bench_n$n.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'clocks',d['clocks'])
for r in d['per_rank_ms']: print(r['rank'], round(r['step_ms'],3), {k:round(v,3) for k,v in r['device'].items()}, round(r['e2e_step_ms'],3), {k:round(v,3) for k,v in r['e2e'].items()})
print(d['strong_scaling_one_view'])
"
done
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q -s 2>&1 | grep -E "nccl|passed|failed|DIST" | tail -5
