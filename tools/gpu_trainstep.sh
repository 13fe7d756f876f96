#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-ts}
timeout 600 python tools/bench_training_step.py > gpurun_out/${TAG}_bench_training_step.json 2> gpurun_out/${TAG}_bench_training_step.err; echo "rc=$?"; tail -1 gpurun_out/${TAG}_bench_training_step.json; tail -3 gpurun_out/${TAG}_bench_training_step.err
BWD_STEPS=1 BWD_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_trainstep_launches.csv \
  python tools/bench_training_step.py > gpurun_out/${TAG}_trainstep_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_trainstep_launches.csv 24
