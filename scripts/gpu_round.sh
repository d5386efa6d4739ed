#!/bin/bash
# Full GPU pass: all -m gpu tests, smoke, op microbench, VAE bench, bench.py.  Usage: gpu_round.sh <tag>
TAG=${1:-r01b}
mkdir -p gpurun_out
echo "=== pytest -m gpu" | tee gpurun_out/tests_${TAG}.log
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -25 | tee -a gpurun_out/tests_${TAG}.log
echo "=== smoke" | tee -a gpurun_out/tests_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a gpurun_out/tests_${TAG}.log
echo "=== ops" ; timeout 600 python scripts/bench_ops.py attn 2>&1 | tee gpurun_out/ops_${TAG}.log
echo "=== vae bench"; timeout 900 python scripts/bench_vae.py 2>&1 | tail -5 | tee gpurun_out/vae_bench_${TAG}.log
echo "=== bench"; timeout 1500 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_${TAG}.log
