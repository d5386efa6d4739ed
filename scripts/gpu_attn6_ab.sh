#!/bin/bash
# A/B of attention6.cu (two softmax threads per row) against attention2.cu:  gpurun --timeout 900 -- 'bash scripts/gpu_attn6_ab.sh <tag>'
TAG=${1:-ab6}
mkdir -p gpurun_out
OUT=gpurun_out/attn6_ab_${TAG}.log
: > $OUT
echo "== debug case" | tee -a $OUT
CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/dev/dbg_attn6.py 2>&1 | tail -12 | tee -a $OUT
grep -q "max err" $OUT || exit 1
echo "== tests" | tee -a $OUT
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -k "attention_alternative" -q -x --no-header -p no:cacheprovider 2>&1 | tail -5 | tee -a $OUT
CE_ATTN6_SPLIT=24 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -k "attention_alternative" -q -x --no-header -p no:cacheprovider 2>&1 | tail -3 | tee -a $OUT
CE_ATTN6_SPLIT=0 CE_ATTN6_POLY=2 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -k "attention_alternative" -q -x --no-header -p no:cacheprovider 2>&1 | tail -3 | tee -a $OUT
echo "== attention2 (default)" | tee -a $OUT
CE_ATTN_V2=2 timeout 60 python scripts/bench_ops.py attnself 2>&1 | grep -v mbarrier | tail -2 | tee -a $OUT
for Q in ${SPLITS:-16 24 0}; do for P in ${POLYS:-0 1 2}; do
  echo "== attention6 SPLIT=$Q POLY=$P" | tee -a $OUT
  CE_ATTN_V2=6 CE_ATTN6_SPLIT=$Q CE_ATTN6_POLY=$P timeout 60 python scripts/bench_ops.py attnself 2>&1 | grep -v mbarrier | tail -2 | tee -a $OUT
done; done
for Q in ${SPLITS:-16 24}; do
echo "== event log attention6 SPLIT=$Q POLY=1" | tee -a $OUT
CE_ATTN_V2=6 CE_ATTN6_SPLIT=$Q CE_ATTN6_POLY=1 timeout 100 python scripts/attn_timing.py 2>&1 | grep -v mbarrier | tail -40 | tee -a $OUT
done
