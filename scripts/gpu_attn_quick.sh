#!/bin/bash
TAG=${1:-aq}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -k attention tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.log
for v in 2 0; do echo "== CE_ATTN_V2=$v" | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v python scripts/attn_timing.py 2>&1 | tee -a gpurun_out/ops_${TAG}.log; CE_ATTN_V2=$v timeout 300 python scripts/bench_ops.py attn 2>&1 | tee -a gpurun_out/ops_${TAG}.log; done
echo "== bench v2"; timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-vae 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_v2.log
echo "== bench v1"; CE_ATTN_V2=0 timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-vae 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_v1.log
