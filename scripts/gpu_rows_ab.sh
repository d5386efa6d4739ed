#!/bin/bash
TAG=${1:-rows}
mkdir -p gpurun_out
for o in 2 3; do echo "== CE_ROW_OCC=$o" | tee -a gpurun_out/ops_${TAG}.log; CE_ROW_OCC=$o timeout 300 python scripts/bench_ops.py rows 2>&1 | tee -a gpurun_out/ops_${TAG}.log; done
CE_ROW_OCC=3 timeout 600 python -m pytest tests/test_gpu_ops.py -k "layernorm or rmsnorm" -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3 | tee -a gpurun_out/ops_${TAG}.log
for o in 2 3; do CE_ROW_OCC=$o python bench.py --steps 4 --warmup 3 --no-vae --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_occ$o.json.log | cut -c1-200; done
