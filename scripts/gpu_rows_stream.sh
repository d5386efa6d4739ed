#!/bin/bash
TAG=${1:-rows}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -k "layernorm or rmsnorm" tests/test_gpu_dit.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/tests_${TAG}.log
for o in 1 0; do echo "== CE_ROW_STREAM=$o" | tee -a gpurun_out/ops_${TAG}.log; CE_ROW_STREAM=$o timeout 300 python scripts/bench_ops.py rows 2>&1 | tee -a gpurun_out/ops_${TAG}.log; done
for o in 1 0; do CE_ROW_STREAM=$o python bench.py --steps 4 --warmup 3 --no-vae --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_stream$o.json.log | cut -c1-210; done
