#!/bin/bash
# quick A/B of the self-attention kernel knobs on one B200: bash scripts/gpu_attn_quick.sh <tag>
TAG=${1:-aq}
mkdir -p gpurun_out
for p in 1 3; do
CE_ATTN_POLY=$p timeout 600 python -m pytest tests/test_gpu_ops.py -k attention tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -4 | tee -a gpurun_out/tests_${TAG}.log
done
for p in 0 1 2 3; do echo "== CE_ATTN_POLY=$p" | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_POLY=$p python scripts/attn_timing.py 2>&1 | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_POLY=$p timeout 300 python scripts/bench_ops.py attn 2>&1 | head -2 | tee -a gpurun_out/ops_${TAG}.log; done
