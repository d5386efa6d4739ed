#!/bin/bash
# First-light run on the B200 box: every test file under its own timeout, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python - <<'PY' > gpurun_out/env.txt 2>&1
import torch, os
print(torch.__version__, torch.cuda.is_available(), torch.cuda.get_device_name(0), os.cpu_count())
PY
for t in "tests/test_gpu_ops.py -k linear" "tests/test_gpu_ops.py -k layernorm" "tests/test_gpu_ops.py -k rmsnorm" "tests/test_gpu_ops.py -k attention" "tests/test_gpu_dit.py"; do
  name=$(echo "$t" | tr ' /' '__')
  echo "=== $t" | tee -a gpurun_out/first_light.log
  timeout 600 python -m pytest $t -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -40 | tee -a gpurun_out/first_light.log
done
