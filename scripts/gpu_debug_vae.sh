#!/bin/bash
mkdir -p gpurun_out
CE_VAE_DEBUG=1 timeout 600 python - > gpurun_out/vae_debug.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from oracle import cases
from tests.test_gpu_vae import _build_vae
case = cases.VAE_CASES["wan_5f_64"]
video, z = cases.vae_inputs(case)
m = _build_vae(case)
mu = m.encode(video.cuda()).latent_dist.mode()
torch.cuda.synchronize()
print("mu finite:", torch.isfinite(mu.float()).all().item())
PY
head -80 gpurun_out/vae_debug.log
echo; echo "=== op microbench"; timeout 900 python scripts/bench_ops.py attn gemm rows conv 2>&1 | tee gpurun_out/bench_ops.log
echo "=== attention + dit tests"; timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dit.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15
