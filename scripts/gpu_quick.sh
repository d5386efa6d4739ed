#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/tests_${TAG}.log
timeout 600 python scripts/bench_ops.py attn rows 2>&1 | tee gpurun_out/ops_${TAG}.log
timeout 1500 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-vae 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.log
