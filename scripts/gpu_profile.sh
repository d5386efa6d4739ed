#!/bin/bash
# ncu evidence for profiles/:  gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh <tag>'   then   python scripts/summarize_profiles.py <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
echo "=== launch list of the bench command (gpu__time_duration only; per-launch times are cold-cache and serialised: shares, not absolutes)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-vae --no-cpu-baseline --no-library-bar > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_vae_${TAG}.csv \
    python scripts/bench_vae.py > gpurun_out/ncu_launches_vae_${TAG}.log 2>&1
echo "=== full captures of the dominant kernels at the 14B / 720p shapes"
timeout 900 ncu --set full --clock-control none -k regex:gemm_bf16_2cta_kernel -s 3 -c 2 -o gpurun_out/prof_gemm_${TAG} -f python scripts/bench_ops.py gemm > gpurun_out/ncu_gemm_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention6_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_attn_${TAG} -f python scripts/bench_ops.py attnself > gpurun_out/ncu_attn_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:attention_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_xattn_${TAG} -f python scripts/bench_ops.py attncross > gpurun_out/ncu_xattn_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"layernorm_stats_kernel|rmsnorm_rope_stats_kernel" -s 8 -c 2 -o gpurun_out/prof_rows_${TAG} -f python bench.py --steps 1 --warmup 1 --no-vae --no-cpu-baseline --no-library-bar > gpurun_out/ncu_rows_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv3d_cl_kernel -s 2 -c 1 -o gpurun_out/prof_conv_${TAG} -f python scripts/bench_ops.py conv > gpurun_out/ncu_conv_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
