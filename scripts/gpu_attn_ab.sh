#!/bin/bash
# A/B of the self-attention variants:  gpurun --timeout 900 -- 'bash scripts/gpu_attn_ab.sh <tag>'
TAG=${1:-ab}
mkdir -p gpurun_out
OUT=gpurun_out/attn_ab_${TAG}.log
: > $OUT
echo "== tests (defaults)" | tee -a $OUT
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -k attention -q -x --no-header -p no:cacheprovider 2>&1 | tail -3 | tee -a $OUT
echo "== tests (SPEC SPIN QUARTERS POLY=2)" | tee -a $OUT
CE_ATTN_SPEC=1 CE_ATTN_SPIN=1 CE_ATTN_QUARTERS=1 CE_ATTN_POLY=2 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -k attention -q -x --no-header -p no:cacheprovider 2>&1 | tail -3 | tee -a $OUT
run() {
  echo "== SPEC=$1 SPIN=$2 QUARTERS=$3 POLY=$4" | tee -a $OUT
  CE_ATTN_SPEC=$1 CE_ATTN_SPIN=$2 CE_ATTN_QUARTERS=$3 CE_ATTN_POLY=$4 timeout 60 python scripts/bench_ops.py attnself 2>&1 | grep -v mbarrier | tail -2 | tee -a $OUT
}
for Q in 0 1; do for N in 0 1; do for P in 0 1 2 3; do run 0 $N $Q $P; done; done; done
for P in 0 1 2; do run 1 1 1 $P; run 1 0 0 $P; done
echo "== timing SPIN=1 QUARTERS=1 POLY=1" | tee -a $OUT
CE_ATTN_SPIN=1 CE_ATTN_QUARTERS=1 CE_ATTN_POLY=1 timeout 60 python scripts/attn_timing.py 2>&1 | grep -v mbarrier | tail -28 | tee -a $OUT
