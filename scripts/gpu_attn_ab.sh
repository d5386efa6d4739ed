#!/bin/bash
# A/B of the self-attention variants:  gpurun --timeout 900 -- 'bash scripts/gpu_attn_ab.sh <tag>'
TAG=${1:-ab}
mkdir -p gpurun_out
OUT=gpurun_out/attn_ab_${TAG}.log
: > $OUT
echo "== tests" | tee -a $OUT
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3 | tee -a $OUT
for FIXED in 1 0; do for SPEC in 0 1; do for POLY in 0 1; do
  echo "== CE_ATTN_FIXED=$FIXED CE_ATTN_SPEC=$SPEC CE_ATTN_POLY=$POLY" | tee -a $OUT
  CE_ATTN_FIXED=$FIXED CE_ATTN_SPEC=$SPEC CE_ATTN_POLY=$POLY timeout 60 python scripts/bench_ops.py attnself 2>&1 | grep -v mbarrier | tail -3 | tee -a $OUT
done; done; done
for SPEC in 0 1; do
  echo "== timing CE_ATTN_FIXED=1 CE_ATTN_SPEC=$SPEC" | tee -a $OUT
  CE_ATTN_FIXED=1 CE_ATTN_SPEC=$SPEC timeout 60 python scripts/attn_timing.py 2>&1 | grep -v mbarrier | tail -32 | tee -a $OUT
done
echo "== cross attention" | tee -a $OUT
timeout 100 python scripts/bench_ops.py attncross 2>&1 | tail -4 | tee -a $OUT
