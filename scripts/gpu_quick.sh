#!/bin/bash
# Fast iteration on one kernel:  gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag> [pytest -k expr] [bench_ops group]'
TAG=${1:-q}; KEXPR=${2:-attention}; OPS=${3:-attn}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -k "$KEXPR" -q -x --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/tests_${TAG}.log
timeout 300 python scripts/bench_ops.py $OPS 2>&1 | tee gpurun_out/ops_${TAG}.log
