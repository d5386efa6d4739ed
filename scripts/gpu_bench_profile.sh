#!/bin/bash
# Bench + profiles on the B200 box.  Usage: gpu_bench_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -k rmsnorm -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/rms_${TAG}.log
echo "=== bench (full 40 layers)" | tee gpurun_out/bench_${TAG}.log
timeout 1200 python bench.py --steps 4 --warmup 3 2>&1 | tail -5 | tee -a gpurun_out/bench_${TAG}.log
echo "=== ncu launch list (2 layers, 1 step)" 
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
   python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "=== ncu full: GEMM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 12 -c 4 -o gpurun_out/prof_gemm_${TAG} -f \
   python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm_${TAG}.log 2>&1
echo "=== ncu full: attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 2 -o gpurun_out/prof_attn_${TAG} -f \
   python bench.py --layers 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn_${TAG}.log 2>&1
ls -la gpurun_out
