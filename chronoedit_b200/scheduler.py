"""`FlowUniPCMultistepScheduler` -- the sampling-loop object either side of the DiT call, with `step` as ONE fused launch.

Mirror of chronoedit/_src/models/fm_solvers_unipc.py (class of the same name; the diffusers CLI's
`UniPCMultistepScheduler(flow_shift=...)` is the same algorithm, scripts/run_inference_diffusers.py:379-382): same
constructor arguments, `set_timesteps`, `step`, `timesteps`, `sigmas`, `model_outputs`, `last_sample`, `order`, so
`ChronoEditPipeline` (chronoedit_diffusers/pipeline_chronoedit.py:667-668, 700-709, 739) and the native loop
(chronoedit_14b_edit_model.py:134-157) drive it unchanged -- including their temporal-reasoning slicing of
`model_outputs` / `last_sample`, which are ordinary tensors here.

What runs where (same split as the reference, which keeps sigmas on the CPU "to avoid too much CPU/GPU communication",
:162, :240):
  * host: the sigma schedule and the handful of fp32 scalars of a step, computed with the same 0-dim fp32 CPU tensor ops
    as the reference (:418-447, :565-620) so that they are bit-identical;
  * device: everything latent-sized -- x0 conversion, corrector, predictor, optionally the classifier-free-guidance combine
    before it and the next model input after it -- in `ce_unipc_step` (csrc/sampler.cu), rounding where the reference's
    separate kernels round.

Only the configuration the reference instantiates is built (solver_order 2, bh2, predict_x0, flow_prediction, no
thresholding, final sigma 0); anything else raises NotImplementedError at construction.  There is no CPU path: `step`
requires CUDA tensors.
"""
from __future__ import annotations

import ctypes
from types import SimpleNamespace
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib

_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


class SchedulerOutput(SimpleNamespace):
    """`prev_sample` holder (diffusers.schedulers.scheduling_utils.SchedulerOutput)."""


class FlowUniPCMultistepScheduler:
    order = 1  # fm_solvers_unipc.py:80

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, prediction_type: str = "flow_prediction",
                 shift: Optional[float] = 1.0, use_dynamic_shifting: bool = False, thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, sample_max_value: float = 1.0, predict_x0: bool = True,
                 solver_type: str = "bh2", lower_order_final: bool = True, disable_corrector: List[int] = [],
                 solver_p=None, timestep_spacing: str = "linspace", steps_offset: int = 0,
                 final_sigmas_type: Optional[str] = "zero"):
        if solver_type in ("midpoint", "heun", "logrho"):   # :111-113
            solver_type = "bh2"
        elif solver_type not in ("bh1", "bh2"):
            raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        unsupported = {
            "solver_order != 2": solver_order != 2, "prediction_type != flow_prediction": prediction_type != "flow_prediction",
            "use_dynamic_shifting": use_dynamic_shifting, "thresholding": thresholding, "predict_x0=False": not predict_x0,
            "solver_type bh1": solver_type != "bh2", "lower_order_final=False": not lower_order_final,
            "disable_corrector": len(disable_corrector) > 0, "solver_p": solver_p is not None,
            "final_sigmas_type != zero": final_sigmas_type != "zero",
        }
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError("chronoedit_b200 builds the scheduler configuration ChronoEdit uses "
                                      "(fm_solvers_unipc.py defaults); not built: " + ", ".join(bad))
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, solver_order=solver_order, prediction_type=prediction_type, shift=shift,
            use_dynamic_shifting=use_dynamic_shifting, thresholding=thresholding,
            dynamic_thresholding_ratio=dynamic_thresholding_ratio, sample_max_value=sample_max_value, predict_x0=predict_x0,
            solver_type=solver_type, lower_order_final=lower_order_final, disable_corrector=list(disable_corrector),
            solver_p=solver_p, timestep_spacing=timestep_spacing, steps_offset=steps_offset, final_sigmas_type=final_sigmas_type)
        self.predict_x0 = predict_x0
        self.num_inference_steps = None
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)   # :127-130
        self.sigmas = sigmas.to("cpu")
        self.timesteps = sigmas * num_train_timesteps
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * solver_order
        self.timestep_list = [None] * solver_order
        self.lower_order_nums = 0
        self.disable_corrector = list(disable_corrector)
        self.solver_p = None
        self.last_sample = None
        self.this_order = 1
        self._step_index = None
        self._begin_index = None
        self._coef_cache = {}
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()

    # ------------------------------------------------------------------ bookkeeping (same surface as the reference)
    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def __len__(self):
        return self.config.num_train_timesteps

    def scale_model_input(self, sample: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return sample

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, sigmas=None, mu=None,
                      shift: Optional[float] = None):
        """Sigma schedule of the run (:174-241): linspace(sigma_max, sigma_min, N+1)[:-1], shifted, a trailing 0; integer
        timesteps by truncation."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        else:
            sigmas = np.asarray(sigmas, dtype=np.float64)
        if shift is None:
            shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.config.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)   # stays on the CPU
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self._timesteps_host = [int(t) for t in self.timesteps.tolist()]
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.timestep_list = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = 1
        self._step_index = None
        self._begin_index = None
        self._coef_cache = {}

    def index_for_timestep(self, timestep, schedule_timesteps=None):   # :646-656, on the host copy (no device sync)
        ts = self._timesteps_host if schedule_timesteps is None else [int(t) for t in schedule_timesteps.tolist()]
        t = int(timestep)
        idx = [i for i, v in enumerate(ts) if v == t]
        return idx[1] if len(idx) > 1 else idx[0]

    def _init_step_index(self, timestep):
        self._step_index = self.index_for_timestep(timestep) if self._begin_index is None else self._begin_index

    # ------------------------------------------------------------------ scalars of one step (host, fp32 0-dim CPU tensors)
    @staticmethod
    def _lambda(sigma: torch.Tensor) -> torch.Tensor:
        return torch.log(1 - sigma) - torch.log(sigma)

    def _bh_scalars(self, i_t: int, i_s0: int):
        """(sigma_t/sigma_s0, alpha_t*h_phi_1, alpha_t*B_h, h, hh, expm1(hh), lambda_s0) for the update sigma[i_s0] -> sigma[i_t]
        (:418-426 / :565-573 with B_h = expm1(hh), :454-455)."""
        s_t, s_0 = self.sigmas[i_t], self.sigmas[i_s0]
        lam_0 = self._lambda(s_0)
        h = self._lambda(s_t) - lam_0
        hh = -h
        e = torch.expm1(hh)
        return (s_t / s_0).item(), ((1 - s_t) * e).item(), ((1 - s_t) * e).item(), h, hh, e, lam_0

    _COEF_FIELDS = ("sigma", "use_corrector", "c_order", "c_inv_rk", "c_rho0", "c_rho1", "c_x", "c_m0", "c_bh", "p_order", "p_x",
                    "p_m0", "p_bh", "p_inv_rk", "p_zero")

    def _fill_coefficients(self, a: "_lib.UniPCStepArgsC", i: int, dtype: torch.dtype):
        """Scalars of step i.  They depend only on (i, order history, dtype), so they are computed once per schedule and cached:
        the steady-state step does no host tensor arithmetic, no 2x2 solve and no .item() beyond this lookup.  (The reference
        builds R, b on the sample's device and solves there, fm_solvers_unipc.py:587-620; here -- as in the oracle that is pinned
        bit for bit against a CPU run of the reference -- the solve runs on the host in fp32, which can differ from a CUDA solve
        in the last ulp of rho before it is rounded to the sample dtype.)"""
        key = (i, self.this_order, i > 0 and self.last_sample is not None, self.lower_order_nums, dtype)
        hit = self._coef_cache.get(key)
        if hit is not None:
            for f, v in zip(self._COEF_FIELDS, hit[0]):
                setattr(a, f, v)
            return hit[1]
        order = self._compute_coefficients(a, i, dtype)
        self._coef_cache[key] = (tuple(getattr(a, f) for f in self._COEF_FIELDS), order)
        return order

    def _compute_coefficients(self, a: "_lib.UniPCStepArgsC", i: int, dtype: torch.dtype):
        a.sigma = self.sigmas[i].item()
        use_corrector = i > 0 and self.last_sample is not None   # :701-705 (disable_corrector is empty)
        a.use_corrector = int(use_corrector)
        a.c_order = self.this_order
        a.c_inv_rk, a.c_rho0, a.c_rho1 = 1.0, 0.0, 0.5
        if use_corrector:
            a.c_x, a.c_m0, a.c_bh, h, hh, e, lam_0 = self._bh_scalars(i, i - 1)
            if self.this_order == 2:   # :575-620
                rk = (self._lambda(self.sigmas[i - 2]) - lam_0) / h
                phi = e / hh - 1
                b1 = phi * 1 / e
                phi = phi / hh - 1 / 2
                b2 = phi * 2 / e
                R = torch.stack([torch.ones(2), torch.stack([rk, torch.tensor(1.0)])])
                rho = torch.linalg.solve(R, torch.stack([b1, b2])).to(dtype)
                a.c_inv_rk = float(np.float32(1.0) / np.float32(rk.item()))   # torch's CUDA tensor/scalar multiplies by 1/b
                a.c_rho0, a.c_rho1 = rho[0].item(), rho[1].item()
        order = min(self.config.solver_order, len(self._timesteps_host) - i, self.lower_order_nums + 1)   # :729-737
        assert order > 0
        a.p_order = order
        a.p_x, a.p_m0, a.p_bh, h, hh, e, lam_0 = self._bh_scalars(i + 1, i)
        a.p_inv_rk, a.p_zero = 1.0, 0.0
        if order == 2:
            rk = (self._lambda(self.sigmas[i - 1]) - lam_0) / h
            a.p_inv_rk = float(np.float32(1.0) / np.float32(rk.item()))
        else:
            a.p_zero = ((1 - self.sigmas[i + 1]) * e * 0).item()
        return order

    # ------------------------------------------------------------------ the step
    def _launch(self, cond, uncond, guidance, timestep, sample, model_input_out) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if (sample.dtype, cond.dtype) not in ((torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16)):
            raise TypeError(f"(sample, model_output) dtypes {sample.dtype}, {cond.dtype} not built: (fp32,fp32), (fp32,bf16), (bf16,bf16)")
        if cond.shape != sample.shape or (uncond is not None and (uncond.shape != sample.shape or uncond.dtype != cond.dtype)):
            raise ValueError("model output and sample must have the same shape")
        if self._step_index is None:
            self._init_step_index(timestep)
        i = self._step_index
        sample, cond = sample.contiguous(), cond.contiguous()
        uncond = None if uncond is None else uncond.contiguous()
        state = [None if t is None else t.contiguous() for t in (self.last_sample, self.model_outputs[-1], self.model_outputs[-2])]
        for t in state:
            if t is not None and (t.shape != sample.shape or t.dtype != sample.dtype):
                raise ValueError(f"scheduler state {tuple(t.shape)} {t.dtype} does not match the sample {tuple(sample.shape)} {sample.dtype} "
                                 "(after a temporal-reasoning cut, slice model_outputs and last_sample as the pipeline does)")
        a = _lib.UniPCStepArgsC()
        order = self._fill_coefficients(a, i, sample.dtype)
        x0 = torch.empty_like(sample)
        prev = torch.empty_like(sample)
        corrected = torch.empty_like(sample) if a.use_corrector else sample
        a.sample_dtype, a.model_dtype, a.n = _DTYPES[sample.dtype], _DTYPES[cond.dtype], sample.numel()
        a.cond, a.uncond, a.guidance = cond.data_ptr(), (uncond.data_ptr() if uncond is not None else None), float(guidance)
        a.sample = sample.data_ptr()
        a.last_sample, a.m_prev, a.m_prev2 = [None if t is None else t.data_ptr() for t in state]
        a.x0_out, a.prev_sample_out = x0.data_ptr(), prev.data_ptr()
        a.corrected_out = corrected.data_ptr() if a.use_corrector else None
        a.model_input_out, a.inner, a.c_lat, a.c_total = None, 1, 1, 1
        if model_input_out is not None:
            if not (model_input_out.dtype == torch.bfloat16 and model_input_out.is_contiguous()
                    and model_input_out.dim() == sample.dim() and model_input_out.shape[0] == sample.shape[0]
                    and model_input_out.shape[2:] == sample.shape[2:] and model_input_out.shape[1] >= sample.shape[1]):
                raise ValueError("model_input_out must be a contiguous bf16 CUDA tensor [B, C_total >= C_latent, T, H, W]")
            a.model_input_out = model_input_out.data_ptr()
            a.inner, a.c_lat, a.c_total = sample[0, 0].numel(), sample.shape[1], model_input_out.shape[1]
        self._native_step(a, dict(cond=cond, uncond=uncond, sample=sample, last_sample=state[0], m_prev=state[1], m_prev2=state[2],
                                  x0=x0, corrected=corrected if a.use_corrector else None, prev=prev, model_input=model_input_out))
        # state update of step() (:720-749)
        self.model_outputs = self.model_outputs[1:] + [x0]
        self.timestep_list = self.timestep_list[1:] + [timestep]
        self.this_order = order
        self.last_sample = corrected
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return prev, x0

    def _native_step(self, a: "_lib.UniPCStepArgsC", tensors: dict) -> None:
        """The one place the scheduler reaches the C ABI.  `a` already holds every pointer and coefficient; `tensors` names the
        same buffers as torch tensors (kept alive across the launch; also what a test harness needs to stand in for the kernel)."""
        if not (tensors["sample"].is_cuda and tensors["cond"].is_cuda):
            raise _lib.CEError("FlowUniPCMultistepScheduler.step needs CUDA tensors: chronoedit_b200 has no CPU path")
        _lib.check(_lib.lib().ce_unipc_step(ctypes.byref(a), _lib.current_stream()))

    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor, return_dict: bool = True,
             generator=None):
        """fm_solvers_unipc.py:670-756.  Returns `(prev_sample, x0_prediction)` or a SchedulerOutput."""
        prev, x0 = self._launch(model_output, None, 0.0, timestep, sample, None)
        if not return_dict:
            return (prev, x0)
        return SchedulerOutput(prev_sample=prev)

    def step_cfg(self, noise_pred: torch.Tensor, noise_uncond: torch.Tensor, guidance_scale: float, timestep, sample: torch.Tensor,
                 model_input_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Fused form of pipeline_chronoedit.py:736-739 (+ :712 for the next iteration): the guidance combine, the scheduler
        step and -- if `model_input_out` [B, 36, T, H, W] bf16 is given -- the latent channels of the next model input, in
        one launch.  Returns prev_sample."""
        return self._launch(noise_pred, noise_uncond, guidance_scale, timestep, sample, model_input_out)[0]

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """:770-812 (not on the inference path; plain torch)."""
        sigmas = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
        if self._begin_index is None:
            idx = [self.index_for_timestep(t) for t in timesteps]
        elif self._step_index is not None:
            idx = [self._step_index] * timesteps.shape[0]
        else:
            idx = [self._begin_index] * timesteps.shape[0]
        sigma = sigmas[idx].flatten()
        while len(sigma.shape) < len(original_samples.shape):
            sigma = sigma.unsqueeze(-1)
        return (1 - sigma) * original_samples + sigma * noise
