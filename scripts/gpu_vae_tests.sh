#!/bin/bash
mkdir -p gpurun_out
for t in "tests/test_gpu_vae.py -k conv" "tests/test_gpu_vae.py -k vae"; do
  echo "=== $t" | tee -a gpurun_out/vae_tests.log
  timeout 900 python -m pytest $t -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -60 | tee -a gpurun_out/vae_tests.log
done
