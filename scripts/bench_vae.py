#!/usr/bin/env python
"""VAE encode/decode timing at the BASELINE geometry (5 px frames, 720x1280), random-init Wan2.1-width weights."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import chronoedit_b200 as ce  # noqa: E402
from oracle import vae_oracle as V  # noqa: E402

torch.manual_seed(0)
cfg = V.VAEConfig.wan21()
m = ce.AutoencoderKLWan()
sd = {k: v.to(torch.bfloat16) for k, v in V.random_state_dict(cfg, 0).items()}
m.load_state_dict(sd)
m = m.cuda()
video = torch.zeros(1, 3, 5, 720, 1280, dtype=torch.bfloat16, device="cuda")
video[:, :, 0] = torch.rand(1, 3, 720, 1280, device="cuda") * 2 - 1
z = torch.randn(1, 16, 2, 90, 160, dtype=torch.bfloat16, device="cuda")


def timeit(fn, iters=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms_e = timeit(lambda: m.encode(video))
le = m.launches()
ms_d = timeit(lambda: m.decode(z))
ld = m.launches()
fe, fd = V.conv_flops(cfg, 5, 720, 1280, False), V.conv_flops(cfg, 5, 720, 1280, True)
print(json.dumps({"op": "vae_encode_720p_5f", "ms": round(ms_e, 2), "conv_tflop": round(fe / 1e12, 2), "tflops": round(fe / ms_e / 1e9, 1), "launches": le,
                  "algorithmic_GB": 15.71, "GBps": round(15.71e9 / ms_e / 1e6, 1)}))
print(json.dumps({"op": "vae_decode_720p_5f", "ms": round(ms_d, 2), "conv_tflop": round(fd / 1e12, 2), "tflops": round(fd / ms_d / 1e9, 1), "launches": ld,
                  "algorithmic_GB": 23.64, "GBps": round(23.64e9 / ms_d / 1e6, 1)}))
print(json.dumps({"workspace_GB": round(m._ws.numel() / 1e9, 2)}))
