#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 0 1 2 4 8 3 6 7 15; do
  SHERF_FRONT_KNOCK=$k timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/knock_$k.json 2> gpurun_out/knock_$k.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/knock_$k.json').read().splitlines() if l.startswith('{')][-1])
s=d['stages_ms_per_view_call']
print('knock $k: front %.3f ms  step %.3f' % (s['front:warp+gather+fusion'], d['ms_per_step']))
"
done
