"""CPU restatement of the per-step sampling glue around the DiT call -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What is restated:
  * the flow-matching UniPC multistep scheduler, `FlowUniPCMultistepScheduler`
    (chronoedit/_src/models/fm_solvers_unipc.py): sigma schedule `set_timesteps` (:174-241), flow-prediction x0 conversion
    (:329-346), UniP predictor bh2 (:365-499), UniC corrector bh2 (:501-641), `step` (:670-756), for the configuration every
    caller in the reference uses: solver_order 2, bh2, predict_x0, flow_prediction, no thresholding, lower_order_final;
  * classifier-free-guidance combine `uncond + g * (cond - uncond)` (chronoedit_diffusers/pipeline_chronoedit.py:736);
  * model-input assembly `cat([latents, condition], dim=1).to(bf16)` (:712).

Every tensor op of the reference rounds its result to the tensor dtype (bf16 latents in the diffusers pipeline,
pipeline_chronoedit.py:676-687; fp32 in the native loop, chronoedit_14b_edit_model.py:121-128), and the scalar coefficients
are fp32 0-dim CPU tensors.  The restatement keeps the same op order and the same rounding points, written as one flat
formula list per step so that the fused CUDA kernel can be checked against it op for op.

CPU vs CUDA torch semantics.  The reference runs on CUDA; the golden vectors are generated on CPU (no GPU in the build
container).  torch evaluates `tensor (op) 0-dim-fp32-CPU-scalar` differently on the two devices (ATen, headers in this image):
  * CPU: TensorIterator first casts the scalar operand to the common dtype (bf16 latents -> the COEFFICIENT is rounded to
    bf16), then the kernel runs;  CUDA: `gpu_kernel_with_scalars` reads the scalar at full fp32 precision
    (ATen/native/cuda/Loops.cuh:185-232, `iter.scalar_value<opmath_t>`), no operand cast;
  * `tensor / scalar`: CPU divides; CUDA multiplies by the fp32 reciprocal (BinaryDivTrueKernel.cu).
The formulas below are written with ordinary torch ops, so they follow the device the tensors live on:
  * on CPU tensors they reproduce the unmodified reference bit for bit (tests/golden/unipc_*, the pin);
  * on CUDA tensors (tests/test_gpu_sampler.py) they are what the reference computes on a GPU -- the fused kernel is required
    to match THAT bit for bit;
  * `cuda_semantics=True` emulates the CUDA behaviour with CPU tensors (explicit fp32 scalar math, reciprocal multiply) for
    places that have no GPU torch run to compare with (smoke()); the GPU tests check the emulation against the real thing.

Pinning: tests/golden/make_golden_unipc.py runs the UNMODIFIED reference file through oracle/diffusers_shim and stores
per-step samples; tests/test_sampler_cpu.py requires bit equality (bf16 and fp32).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np
import torch


@dataclasses.dataclass
class StepCoeffs:
    """Scalars of one scheduler step (all fp32 values held as Python floats)."""
    sigma: float = 0.0          # x0 = x - sigma * v
    use_corrector: bool = False
    c_order: int = 0
    c_x: float = 0.0            # sigma_t / sigma_s0
    c_m0: float = 0.0           # alpha_t * h_phi_1
    c_bh: float = 0.0           # alpha_t * B_h
    c_rk: float = 1.0           # order 2 only
    c_rho0: float = 0.0         # order 2 only (already rounded to the sample dtype)
    c_rho1: float = 0.5
    p_order: int = 1
    p_x: float = 0.0
    p_m0: float = 0.0
    p_bh: float = 0.0
    p_rk: float = 1.0
    p_zero: float = 0.0         # order 1: the scalar `alpha_t * B_h * 0` that is still subtracted (sign of zero)


def flow_sigmas(num_inference_steps: int, shift: float, num_train_timesteps: int = 1000, base_shift: float = 1.0):
    """(sigmas fp32 [N+1] with a trailing 0, timesteps int64 [N])  (fm_solvers_unipc.py:119-134, 196-241)."""
    alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
    train = torch.from_numpy(1.0 - alphas).to(torch.float32)
    train = base_shift * train / (1 + (base_shift - 1) * train)
    s_max, s_min = train[0].item(), train[-1].item()
    sig = np.linspace(s_max, s_min, num_inference_steps + 1).copy()[:-1]
    sig = shift * sig / (1 + (shift - 1) * sig)
    timesteps = torch.from_numpy(sig * num_train_timesteps).to(torch.int64)
    sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))
    return sigmas, timesteps


def _lam(s):
    return torch.log(1 - s) - torch.log(s)


def _bh(sigmas, i_t, i_s0):
    s_t, s_0 = sigmas[i_t], sigmas[i_s0]
    h = _lam(s_t) - _lam(s_0)
    hh = -h
    e = torch.expm1(hh)
    return s_t, s_0, h, hh, e


def step_coeffs(sigmas: torch.Tensor, i: int, n_steps: int, lower_order_nums: int, prev_order: int, have_last: bool,
                dtype: torch.dtype) -> StepCoeffs:
    c = StepCoeffs(sigma=sigmas[i].item())
    if i > 0 and have_last:   # UniC at the current sigma, from sigma[i-1]  (:565-640)
        s_t, s_0, h, hh, e = _bh(sigmas, i, i - 1)
        c.use_corrector, c.c_order = True, prev_order
        c.c_x, c.c_m0, c.c_bh = (s_t / s_0).item(), ((1 - s_t) * e).item(), ((1 - s_t) * e).item()
        if prev_order == 2:
            rk = (_lam(sigmas[i - 2]) - _lam(s_0)) / h
            phi = e / hh - 1
            b1 = phi * 1 / e
            phi = phi / hh - 1 / 2
            b2 = phi * 2 / e
            R = torch.stack([torch.ones(2), torch.stack([rk, torch.tensor(1.0)])])
            rho = torch.linalg.solve(R, torch.stack([b1, b2])).to(dtype)
            c.c_rk, c.c_rho0, c.c_rho1 = rk.item(), rho[0].item(), rho[1].item()
    order = min(2, n_steps - i, lower_order_nums + 1)   # (:729-737)
    s_t, s_0, h, hh, e = _bh(sigmas, i + 1, i)
    c.p_order = order
    c.p_x, c.p_m0, c.p_bh = (s_t / s_0).item(), ((1 - s_t) * e).item(), ((1 - s_t) * e).item()
    if order == 2:
        c.p_rk = ((_lam(sigmas[i - 1]) - _lam(s_0)) / h).item()
    else:
        c.p_zero = ((1 - s_t) * e * 0).item()
    return c


def _t(v: float) -> torch.Tensor:   # fp32 0-dim CPU tensor: scalar operand that does not take part in type promotion
    return torch.tensor(v, dtype=torch.float32)


class _Ops:
    """tensor-with-scalar ops either as torch evaluates them on the tensor's device (`emulate_cuda=False`) or with CUDA's
    scalar handling spelled out on CPU tensors (`emulate_cuda=True`)."""

    def __init__(self, emulate_cuda: bool):
        self.emu = emulate_cuda

    def smul(self, s: float, t: torch.Tensor) -> torch.Tensor:
        if self.emu:
            return (t.float() * _t(s)).to(t.dtype)
        return _t(s) * t

    def ssub(self, t: torch.Tensor, s: float) -> torch.Tensor:
        if self.emu:
            return (t.float() - _t(s)).to(t.dtype)
        return t - _t(s)

    def sdiv(self, t: torch.Tensor, s: float) -> torch.Tensor:
        if self.emu:
            return (t.float() * _t(float(np.float32(1.0) / np.float32(s)))).to(t.dtype)
        return t / _t(s)

    def pymul(self, g: float, t: torch.Tensor) -> torch.Tensor:   # Python-number operand (guidance scale)
        if self.emu:
            return (t.float() * _t(float(np.float32(g)))).to(t.dtype)
        return g * t


def cfg_combine(cond: torch.Tensor, uncond: torch.Tensor, guidance: float, cuda_semantics: bool = False) -> torch.Tensor:
    return uncond + _Ops(cuda_semantics).pymul(guidance, cond - uncond)   # pipeline_chronoedit.py:736


def step_formulas(c: StepCoeffs, v, x, last, m_prev, m_prev2, cuda_semantics: bool = False):
    """One scheduler step as a flat formula list -> (x0 prediction, corrected sample, next sample)."""
    o = _Ops(cuda_semantics)
    dev = x.device
    m_t = x - o.smul(c.sigma, v)
    if c.use_corrector:
        xt_ = o.smul(c.c_x, last) - o.smul(c.c_m0, m_prev)
        d_t = m_t - m_prev
        if c.c_order == 1:
            inner = 0 + torch.tensor(0.5, dtype=last.dtype, device=dev) * d_t
        else:
            d1 = o.sdiv(m_prev2 - m_prev, c.c_rk)
            inner = torch.tensor(c.c_rho0, dtype=last.dtype, device=dev) * d1 + torch.tensor(c.c_rho1, dtype=last.dtype, device=dev) * d_t
        x = (xt_ - o.smul(c.c_bh, inner)).to(last.dtype)
    xt_ = o.smul(c.p_x, x) - o.smul(c.p_m0, m_t)
    if c.p_order == 1:
        nxt = o.ssub(xt_, c.p_zero)
    else:
        d1 = o.sdiv(m_prev - m_t, c.p_rk)
        nxt = xt_ - o.smul(c.p_bh, torch.tensor(0.5, dtype=x.dtype, device=dev) * d1)
    return m_t, x, nxt.to(x.dtype)


class UniPCOracle:
    """Stateful wrapper with the reference object's surface: set_timesteps / step / model_outputs / last_sample."""

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, cuda_semantics: bool = False):
        self.num_train_timesteps, self.base_shift, self.cuda_semantics = num_train_timesteps, shift, cuda_semantics
        self.model_outputs: List[Optional[torch.Tensor]] = [None, None]
        self.last_sample = None

    def set_timesteps(self, num_inference_steps: int, shift: Optional[float] = None):
        self.sigmas, self.timesteps = flow_sigmas(num_inference_steps, self.base_shift if shift is None else shift,
                                                  self.num_train_timesteps, self.base_shift)
        self.n = num_inference_steps
        self.model_outputs, self.last_sample = [None, None], None
        self.lower_order_nums, self.step_index, self.this_order = 0, 0, 1

    def step(self, v: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        i = self.step_index
        c = step_coeffs(self.sigmas, i, self.n, self.lower_order_nums, self.this_order, self.last_sample is not None, sample.dtype)
        m_t, x, nxt = step_formulas(c, v, sample, self.last_sample, self.model_outputs[1], self.model_outputs[0], self.cuda_semantics)
        self.model_outputs = [self.model_outputs[1], m_t]
        self.this_order, self.last_sample = c.p_order, x
        self.lower_order_nums = min(self.lower_order_nums + 1, 2)
        self.step_index += 1
        return nxt

    def cut_frames(self, keep=(0, -1)):
        """Temporal-reasoning cut of the scheduler state (pipeline_chronoedit.py:700-709)."""
        idx = list(keep)
        self.model_outputs = [None if m is None else m[:, :, idx] for m in self.model_outputs]
        if self.last_sample is not None:
            self.last_sample = self.last_sample[:, :, idx]


def model_input(latents: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
    return torch.cat([latents, condition], dim=1).to(torch.bfloat16)   # pipeline_chronoedit.py:712
