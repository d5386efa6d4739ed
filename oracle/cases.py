"""Seeded parity cases shared by the golden generator, the tests, smoke() and bench.py.

TEST INFRASTRUCTURE ONLY.  Weights and inputs are regenerated from seeds on the torch CPU
generator (deterministic for a given torch build; the fixtures carry checksums so that RNG
drift is detected rather than silently compared against stale outputs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import dit_oracle as D
from . import vae_oracle as V


@dataclass
class DiTCase:
    name: str
    cfg: D.DiTConfig
    batch: int
    frames: int          # latent frames
    height: int          # latent height
    width: int           # latent width
    timestep: int
    text_len: int = 512
    text_valid: int = 100  # rows >= text_valid are zero (pipeline_chronoedit.py:234-237 zero-pads prompts)
    weight_seed: int = 0
    input_seed: int = 1234


# BASELINE.json configs[0]: 2-layer / dim-256 random-init DiT (plumbing / correctness)
DIT_CASES: Dict[str, DiTCase] = {
    c.name: c
    for c in [
        DiTCase("tiny_t2", D.DiTConfig.tiny(), batch=1, frames=2, height=16, width=16, timestep=749),
        DiTCase("tiny_t8", D.DiTConfig.tiny(), batch=1, frames=8, height=8, width=12, timestep=499),
        DiTCase("tiny_b2", D.DiTConfig.tiny(), batch=2, frames=2, height=12, width=20, timestep=999),
        # ragged: L = 2*9*13 = 234 tokens (not a multiple of any tile), heads = 3
        DiTCase("tiny_ragged", D.DiTConfig.tiny(heads=3, ffn_dim=640), batch=1, frames=2, height=18, width=26,
                timestep=3, text_valid=7),
        # "64x64x5 latent" of configs[0]: hidden_states [1,36,2,64,64] (5 pixel frames at 512x512), L = 2048
        DiTCase("cfg0_64x64", D.DiTConfig.tiny(), batch=1, frames=2, height=64, width=64, timestep=249),
    ]
}


def dit_inputs(case: DiTCase, dtype: torch.dtype = torch.float32):
    g = torch.Generator(device="cpu").manual_seed(case.input_seed)
    cfg = case.cfg
    x = torch.randn(case.batch, cfg.in_channels, case.frames, case.height, case.width, generator=g)
    text = torch.randn(case.batch, case.text_len, cfg.text_dim, generator=g)
    text[:, case.text_valid:] = 0
    img = torch.randn(case.batch, 257, cfg.image_dim, generator=g)
    t = torch.full((case.batch,), case.timestep, dtype=torch.int64)
    return x.to(dtype), t, text.to(dtype), img.to(dtype)


def dit_weights(case: DiTCase, dtype: torch.dtype = torch.float32):
    return D.random_state_dict(case.cfg, seed=case.weight_seed, dtype=dtype)


@dataclass
class VAECase:
    name: str
    cfg: V.VAEConfig
    frames_px: int
    height: int
    width: int
    weight_seed: int = 0
    input_seed: int = 4321


VAE_CASES: Dict[str, VAECase] = {
    c.name: c
    for c in [
        VAECase("tiny_5f", V.VAEConfig.tiny(32), frames_px=5, height=32, width=48),
        VAECase("tiny_9f", V.VAEConfig.tiny(32), frames_px=9, height=48, width=32),
        VAECase("tiny_1f", V.VAEConfig.tiny(32), frames_px=1, height=40, width=24),
        # full-width channels (96..384) at a small geometry: exercises the real tile shapes
        VAECase("wan_5f_64", V.VAEConfig.wan21(), frames_px=5, height=64, width=64),
    ]
}


def vae_inputs(case: VAECase, dtype: torch.dtype = torch.float32):
    """video: frame 0 = image in [-1,1], later frames 0 (pipeline_chronoedit.py:421-425); z ~ N(0,1)."""
    g = torch.Generator(device="cpu").manual_seed(case.input_seed)
    video = torch.zeros(1, 3, case.frames_px, case.height, case.width)
    video[:, :, 0] = torch.rand(1, 3, case.height, case.width, generator=g) * 2 - 1
    if case.frames_px > 1:  # a moving second case so temporal paths see non-zero data
        video[:, :, 1:] = 0.25 * (torch.rand(1, 3, case.frames_px - 1, case.height, case.width, generator=g) * 2 - 1)
    tl = 1 + (case.frames_px - 1) // 4
    z = torch.randn(1, case.cfg.z_dim, tl, case.height // 8, case.width // 8, generator=g)
    return video.to(dtype), z.to(dtype)


def vae_weights(case: VAECase, dtype: torch.dtype = torch.float32):
    return V.random_state_dict(case.cfg, seed=case.weight_seed, dtype=dtype)


def to_bf16_state(sd: Dict[str, torch.Tensor], keep_fp32=D.KEEP_FP32) -> Dict[str, torch.Tensor]:
    """Cast a fp32 state dict the way `from_pretrained(torch_dtype=bf16)` does: everything bf16 except the
    modules listed in `_keep_in_fp32_modules` (transformer_chronoedit.py:338)."""
    return {k: (v if any(s in k for s in keep_fp32) else v.to(torch.bfloat16)) for k, v in sd.items()}


def checksum(t: torch.Tensor) -> float:
    """Order-sensitive fingerprint of a tensor (fp64 weighted sum)."""
    f = t.detach().double().flatten()
    w = torch.arange(1, f.numel() + 1, dtype=torch.float64).remainder(1009.0) + 1.0
    return float((f * w).sum())
