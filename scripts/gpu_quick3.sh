#!/bin/bash
TAG=${1:-q3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/tests_${TAG}.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee -a gpurun_out/tests_${TAG}.log
python bench.py --steps 4 --warmup 3 --no-vae --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.json.log
