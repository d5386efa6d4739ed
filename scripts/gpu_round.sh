#!/bin/bash
# What the driver runs at round end, in one gpurun call:  gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-round}
mkdir -p gpurun_out
echo "=== tests" | tee gpurun_out/tests_${TAG}.log
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee -a gpurun_out/tests_${TAG}.log
echo "=== smoke" | tee -a gpurun_out/tests_${TAG}.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee -a gpurun_out/tests_${TAG}.log
echo "=== ops" | tee gpurun_out/ops_${TAG}.log
timeout 600 python scripts/bench_ops.py attn attnlib gemm rows conv 2>&1 | tee -a gpurun_out/ops_${TAG}.log
timeout 300 python scripts/bench_vae.py 2>&1 | tail -3 | tee -a gpurun_out/ops_${TAG}.log
echo "=== bench"; timeout 1500 python bench.py --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.json.log
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_${TAG}.json.log
