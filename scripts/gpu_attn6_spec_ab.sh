#!/bin/bash
# attention6.cu: speculative first batch of exponentials (CE_ATTN6_SPEC) A/B
TAG=${1:-spec}
mkdir -p gpurun_out
OUT=gpurun_out/attn6_spec_ab_${TAG}.log
: > $OUT
echo "== tests CE_ATTN6_SPEC=1" | tee -a $OUT
CE_ATTN6_SPEC=1 timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -k "attention" -q -x --no-header -p no:cacheprovider 2>&1 | tail -4 | tee -a $OUT
for rep in 1 2; do for X in 0 1; do for P in 1; do
  echo "== SPEC=$X POLY=$P" | tee -a $OUT
  CE_ATTN6_SPEC=$X CE_ATTN6_POLY=$P timeout 60 python scripts/bench_ops.py attnself 2>&1 | grep -v mbarrier | tail -2 | tee -a $OUT
done; done; done
echo "== event log SPEC=1" | tee -a $OUT
CE_ATTN_V2=6 CE_ATTN6_SPEC=1 timeout 100 python scripts/attn_timing.py 2>&1 | grep -v mbarrier | tail -40 | tee -a $OUT
