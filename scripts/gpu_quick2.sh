#!/bin/bash
TAG=${1:-q2}
mkdir -p gpurun_out
python scripts/bench_ops.py attn 2>&1 | tee gpurun_out/ops_${TAG}.log
python bench.py --steps 4 --warmup 3 --no-vae --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.json.log
