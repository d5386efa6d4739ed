#!/bin/bash
TAG=${1:-aq}
mkdir -p gpurun_out
CE_ATTN_V2=3 timeout 600 python -m pytest tests/test_gpu_ops.py -k attention tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/tests_${TAG}.log
for v in 3 2; do echo "== CE_ATTN_V2=$v" | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v python scripts/attn_timing.py 2>&1 | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v timeout 300 python scripts/bench_ops.py attn 2>&1 | tee -a gpurun_out/ops_${TAG}.log; done
