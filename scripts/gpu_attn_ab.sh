#!/bin/bash
# A/B of the self-attention kernels (CE_ATTN_V2 = 4 cluster kernel vs 2):  gpurun --timeout 900 -- 'bash scripts/gpu_attn_ab.sh <tag> [quick]'
TAG=${1:-ab}
mkdir -p gpurun_out
if [ "$2" != "quick" ]; then
CE_ATTN_V2=4 timeout 600 python -m pytest tests/test_gpu_ops.py -k attention tests/test_gpu_dit.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/tests_${TAG}.log
fi
for v in 4 2; do echo "== CE_ATTN_V2=$v" | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v timeout 120 python scripts/attn_timing.py 2>&1 | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v timeout 300 python scripts/bench_ops.py attn 2>&1 | head -2 | tee -a gpurun_out/ops_${TAG}.log; done
