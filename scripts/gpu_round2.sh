#!/bin/bash
TAG=${1:-r01d}
mkdir -p gpurun_out
echo "=== tests (2-CTA GEMM on)" | tee gpurun_out/tests_${TAG}.log
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee -a gpurun_out/tests_${TAG}.log
echo "=== ops: gemm 2cta" | tee gpurun_out/ops_${TAG}.log
timeout 600 python scripts/bench_ops.py gemm rows 2>&1 | tee -a gpurun_out/ops_${TAG}.log
echo "=== ops: gemm 1cta" | tee -a gpurun_out/ops_${TAG}.log
CE_GEMM_2CTA=0 timeout 600 python scripts/bench_ops.py gemm 2>&1 | grep '"gemm"' | tee -a gpurun_out/ops_${TAG}.log
echo "=== bench 2cta"; timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.log
echo "=== bench 1cta"; CE_GEMM_2CTA=0 timeout 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vae 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_1cta.log
