#!/bin/bash
# A/B on the same box: $1 = tag, $2 = env var to toggle (set to 0 for the B arm)
TAG=${1:-ab}; VAR=${2:-CE_ATTN_V2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/tests_${TAG}.log
echo "== A (default)" | tee gpurun_out/ops_${TAG}.log
timeout 600 python scripts/bench_ops.py attn 2>&1 | tee -a gpurun_out/ops_${TAG}.log
echo "== B ($VAR=0)" | tee -a gpurun_out/ops_${TAG}.log
env $VAR=0 timeout 600 python scripts/bench_ops.py attn 2>&1 | tee -a gpurun_out/ops_${TAG}.log
echo "== bench A"; timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-vae 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_A.log
echo "== bench B"; env $VAR=0 timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-vae 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_B.log
