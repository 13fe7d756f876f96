#!/bin/bash
# One gpurun call that refreshes the final round-2 evidence under gpurun_out/ (copied to profiles/ by hand): full GPU test log, the default
# bench line (with the reference CPU arm, PSNR and the training_step key), importance / reference lines, the training benches, the sparse
# encoder timing, the ncu launch lists of the forward bench and of one training step, and --set full captures of the dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2f}
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/bench_brief.py c2 < gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --importance 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_importance64.json 2>> gpurun_out/${TAG}_bench.err
python tools/bench_brief.py imp64 < gpurun_out/${TAG}_bench_importance64.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 300 python tools/bench_backward.py > gpurun_out/${TAG}_bench_backward.json 2>> gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench_backward.json | cut -c1-300
SHERF_BWD_SIMT=1 timeout 300 python tools/bench_backward.py > gpurun_out/${TAG}_bench_backward_fp32_fma.json 2>> gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench_backward_fp32_fma.json | cut -c1-200
timeout 300 python tools/bench_training_step.py > gpurun_out/${TAG}_bench_training_step.json 2>> gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench_training_step.json | cut -c1-300
python tools/time_sparse_encoder.py > gpurun_out/${TAG}_sparse_encoder.json 2>> gpurun_out/${TAG}_bench.err; SHERF_SP_SIMT=1 python tools/time_sparse_encoder.py >> gpurun_out/${TAG}_sparse_encoder.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_sparse_encoder.json | cut -c1-120
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-training-step > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_launches.csv 10
BWD_STEPS=1 BWD_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_trainstep_launches.csv \
  python tools/bench_training_step.py > gpurun_out/${TAG}_trainstep_launches.log 2>&1
python tools/launch_shares.py gpurun_out/${TAG}_trainstep_launches.csv 16
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_front_fused|^k_xformer_bf16|k_decoder_pp|k_cull_search|k_cull_candidates' -s 0 -c 5 -f \
  -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-training-step > gpurun_out/${TAG}_full.log 2>&1
tail -1 gpurun_out/${TAG}_full.log | cut -c1-200
BWD_STEPS=1 BWD_WARMUP=0 timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_umma_grad_w|k_umma_linear' -s 22 -c 2 -f \
  -o gpurun_out/${TAG}_bwd_full python tools/bench_backward.py > gpurun_out/${TAG}_bwd_full.log 2>&1
tail -1 gpurun_out/${TAG}_bwd_full.log | cut -c1-200
