#!/bin/bash
TAG=${1:-poll}
mkdir -p gpurun_out
for v in 2 5; do for ns in 0 20 60 150; do echo "== CE_ATTN_V2=$v CE_ATTN_POLL_NS=$ns" | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v CE_ATTN_POLL_NS=$ns timeout 300 python scripts/bench_ops.py attn 2>&1 | sed -n '1p;3p;5p' | tee -a gpurun_out/ops_${TAG}.log; done; done
CE_ATTN_V2=5 CE_ATTN_POLL_NS=60 timeout 120 python scripts/attn_timing.py 2>&1 | tee -a gpurun_out/ops_${TAG}.log
