#!/bin/bash
TAG=${1:-r01c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/tests_${TAG}.log
timeout 300 python scripts/bench_ops.py rows 2>&1 | tee gpurun_out/ops_${TAG}.log
echo "=== ncu launch list: VAE decode+encode 720p"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_vae_${TAG}.csv python scripts/bench_vae.py > gpurun_out/ncu_vae_${TAG}.log 2>&1
echo "=== ncu full: conv (96ch 720p, 192ch, 384ch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3d_cl_kernel -s 2 -c 3 -o gpurun_out/prof_conv_${TAG} -f python scripts/bench_ops.py conv > gpurun_out/ncu_conv_${TAG}.log 2>&1
echo "=== ncu full: attention (new)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn_${TAG} -f python scripts/bench_ops.py attn > gpurun_out/ncu_attn_${TAG}.log 2>&1
echo "=== bench"; timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.log
ls -la gpurun_out | tail -12
