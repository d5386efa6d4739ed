"""Seeded cases for the sampling-glue oracle / kernel (TEST INFRASTRUCTURE ONLY).  Inputs are regenerated from the seed on
both sides (container and GPU box); only the reference outputs live in tests/golden/unipc_*.safetensors."""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import torch


@dataclasses.dataclass(frozen=True)
class UniPCCase:
    steps: int
    shift: float
    shape: Tuple[int, int, int, int, int]     # [B, 16, Tl, H, W] latent
    sample_dtype: torch.dtype
    model_dtype: torch.dtype
    guidance: Optional[float] = None          # classifier-free guidance scale; None = single forward
    cut_at: Optional[int] = None              # temporal-reasoning cut before this step (frames [0, -1] kept)
    seed: int = 0


UNIPC_CASES = {
    # diffusers pipeline: bf16 latents and model outputs, distilled 8-step schedule (README.md:118 flow_shift 2.0)
    "bf16_8step": UniPCCase(8, 2.0, (1, 16, 2, 6, 10), torch.bfloat16, torch.bfloat16, guidance=None, seed=11),
    # ... with classifier-free guidance 5.0 and the default shift (run_inference_diffusers.py:203-207)
    "bf16_cfg_10step": UniPCCase(10, 5.0, (2, 16, 2, 4, 6), torch.bfloat16, torch.bfloat16, guidance=5.0, seed=12),
    # native loop: fp32 latents (chronoedit_14b_edit_model.py:121-128)
    "fp32_10step": UniPCCase(10, 5.0, (1, 16, 3, 4, 6), torch.float32, torch.float32, seed=13),
    "fp32_bf16model_6step": UniPCCase(6, 3.0, (1, 16, 2, 4, 6), torch.float32, torch.bfloat16, seed=14),
    # temporal reasoning: 4 latent frames until step 3, then frames [0, -1] (pipeline_chronoedit.py:700-709)
    "bf16_cut_8step": UniPCCase(8, 5.0, (1, 16, 4, 4, 6), torch.bfloat16, torch.bfloat16, guidance=3.0, cut_at=3, seed=15),
    # lower_order_final edge cases
    "bf16_1step": UniPCCase(1, 2.0, (1, 16, 2, 4, 6), torch.bfloat16, torch.bfloat16, seed=16),
    "fp32_2step": UniPCCase(2, 2.0, (1, 16, 2, 4, 6), torch.float32, torch.float32, guidance=2.0, seed=17),
}


def case_inputs(case: UniPCCase):
    """(initial sample, [cond_i], [uncond_i]) -- full frame count for every step (callers slice after the cut)."""
    g = torch.Generator().manual_seed(case.seed)
    x = torch.randn(case.shape, generator=g, dtype=torch.float32).to(case.sample_dtype)
    cond = [torch.randn(case.shape, generator=g, dtype=torch.float32).to(case.model_dtype) for _ in range(case.steps)]
    uncond = [torch.randn(case.shape, generator=g, dtype=torch.float32).to(case.model_dtype) for _ in range(case.steps)]
    return x, cond, uncond
