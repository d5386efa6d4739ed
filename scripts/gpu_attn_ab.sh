#!/bin/bash
# A/B of the self-attention kernels:  gpurun --timeout 900 -- 'bash scripts/gpu_attn_ab.sh <tag> "<versions, first one is tested>" [quick]'
TAG=${1:-ab}; VERS=${2:-"5 2"}
mkdir -p gpurun_out
FIRST=${VERS%% *}
if [ "$3" != "quick" ]; then
CE_ATTN_V2=$FIRST timeout 600 python -m pytest tests/test_gpu_ops.py -k attention tests/test_gpu_dit.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/tests_${TAG}.log
fi
for v in $VERS; do echo "== CE_ATTN_V2=$v" | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v timeout 120 python scripts/attn_timing.py 2>&1 | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v timeout 300 python scripts/bench_ops.py attn 2>&1 | head -2 | tee -a gpurun_out/ops_${TAG}.log; done
