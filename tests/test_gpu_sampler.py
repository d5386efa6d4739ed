"""GPU parity tests (-m gpu) of the fused sampling-glue launch (csrc/sampler.cu through ce_unipc_step and the
`FlowUniPCMultistepScheduler` mirror).

The expected values are the oracle's formula list -- pinned bit for bit to the unmodified reference on CPU
(tests/golden/unipc_*) -- evaluated by torch on CUDA tensors, i.e. the reference's own op sequence under torch's CUDA
semantics (fp32 scalar operands at full precision, tensor/scalar as a reciprocal multiply; oracle/unipc_oracle.py header).
The fused kernel must match that bit for bit on every step of every case and at the full 720p latent size.  Against the
golden vectors themselves (a CPU run of the reference, where torch rounds the coefficients to bf16 first) the distance is
bounded by what the manifest recorded for that semantic difference."""
import json
import os

import pytest
import torch
from safetensors.torch import load_file

from oracle import unipc_oracle
from oracle.unipc_cases import UNIPC_CASES, UniPCCase, case_inputs

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32)


def _run_oracle(case, device="cuda", emulate=False):
    """device="cuda": torch ops on CUDA tensors (native CUDA semantics); device="cpu", emulate=True: the CPU emulation of them."""
    o = unipc_oracle.UniPCOracle(shift=1.0, cuda_semantics=emulate)
    o.set_timesteps(case.steps, shift=case.shift)
    x, cond, uncond = case_inputs(case)
    x, cond, uncond = x.to(device), [c.to(device) for c in cond], [u.to(device) for u in uncond]
    outs, x0s = [], []
    for i in range(case.steps):
        if case.cut_at is not None and i == case.cut_at:
            x = x[:, :, [0, -1]]
            o.cut_frames()
        c, u = cond[i], uncond[i]
        if case.cut_at is not None and i >= case.cut_at:
            c, u = c[:, :, [0, -1]], u[:, :, [0, -1]]
        v = unipc_oracle.cfg_combine(c, u, case.guidance, emulate) if case.guidance is not None else c
        x = o.step(v, x)
        outs.append(x.cpu())
        x0s.append(o.model_outputs[-1].cpu())
    return o, outs, x0s


def _run_mirror(case, fused_cfg):
    """Drive the mirror exactly as pipeline_chronoedit.py:693-739 does (slicing included)."""
    from chronoedit_b200.scheduler import FlowUniPCMultistepScheduler
    s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(case.steps, device="cuda", shift=case.shift)
    x, cond, uncond = case_inputs(case)
    x = x.cuda()
    outs, x0s = [], []
    for i, t in enumerate(s.timesteps):
        if case.cut_at is not None and i == case.cut_at:
            x = x[:, :, [0, -1]]
            for j in range(len(s.model_outputs)):
                if s.model_outputs[j] is not None and x.shape[-3] != s.model_outputs[j].shape[-3]:
                    s.model_outputs[j] = s.model_outputs[j][:, :, [0, -1]]
            if s.last_sample is not None:
                s.last_sample = s.last_sample[:, :, [0, -1]]
        c, u = cond[i].cuda(), uncond[i].cuda()
        if case.cut_at is not None and i >= case.cut_at:
            c, u = c[:, :, [0, -1]], u[:, :, [0, -1]]
        if case.guidance is None:
            x = s.step(c, t, x, return_dict=False)[0]
        elif fused_cfg:
            x = s.step_cfg(c, u, case.guidance, t, x)
        else:
            x = s.step(u + case.guidance * (c - u), t, x, return_dict=False)[0]
        outs.append(x)
        x0s.append(s.model_outputs[-1])
    return s, outs, x0s


@pytest.mark.parametrize("fused_cfg", [False, True])
@pytest.mark.parametrize("name", sorted(UNIPC_CASES))
def test_step_bit_exact_vs_oracle_and_golden(name, fused_cfg, golden_dir):
    case = UNIPC_CASES[name]
    if fused_cfg and case.guidance is None:
        pytest.skip("no guidance in this case")
    _, ora, ora_x0 = _run_oracle(case, "cuda")
    s, got, got_x0 = _run_mirror(case, fused_cfg)
    gold = load_file(os.path.join(golden_dir, f"unipc_{name}.safetensors"))
    man = json.load(open(os.path.join(golden_dir, "UNIPC_MANIFEST.json")))["cases"][name]
    for i in range(case.steps):
        assert got[i].dtype == ora[i].dtype and got[i].shape == ora[i].shape
        assert torch.equal(_bits(got[i].cpu()), _bits(ora[i])), f"{name}: sample after step {i} differs from the oracle"
        assert torch.equal(_bits(got_x0[i].cpu()), _bits(ora_x0[i])), f"{name}: x0 prediction of step {i} differs from the oracle"
        g = gold[f"step{i:02d}"]
        if man["elements_changed_by_cuda_semantics"] == 0:
            assert torch.equal(_bits(got[i].cpu()), _bits(g)), f"{name}: step {i} differs from the unmodified reference"
        else:   # the golden vectors are a CPU run: coefficient rounding / true division differ (recorded in the manifest)
            assert (got[i].cpu().float() - g.float()).abs().max().item() <= man["max_abs_change_by_cuda_semantics"] * 1.0001 + 1e-12
    assert s.step_index == case.steps and s.lower_order_nums == min(case.steps, 2)


@pytest.mark.parametrize("name", sorted(UNIPC_CASES))
def test_cpu_emulation_of_cuda_semantics_is_exact(name):
    """The oracle's `cuda_semantics=True` mode (used by smoke() and by the manifest's distance numbers) spells out what torch
    does on CUDA; here it is checked against torch on CUDA itself."""
    case = UNIPC_CASES[name]
    _, native, native_x0 = _run_oracle(case, "cuda")
    _, emu, emu_x0 = _run_oracle(case, "cpu", emulate=True)
    for i in range(case.steps):
        assert torch.equal(_bits(native[i]), _bits(emu[i])) and torch.equal(_bits(native_x0[i]), _bits(emu_x0[i])), f"{name}: step {i}"


def test_full_size_cfg_step_with_model_input():
    """720p / 5-frame latent [1,16,2,90,160], bf16, CFG 5.0, 6 steps; the fused launch also writes the latent channels of the
    next model input (pipeline_chronoedit.py:712)."""
    from chronoedit_b200.scheduler import FlowUniPCMultistepScheduler
    case = UniPCCase(6, 5.0, (1, 16, 2, 90, 160), torch.bfloat16, torch.bfloat16, guidance=5.0, seed=21)
    _, ora, _ = _run_oracle(case, "cuda")
    s = FlowUniPCMultistepScheduler(shift=1)
    s.set_timesteps(case.steps, device="cuda", shift=case.shift)
    x, cond, uncond = case_inputs(case)
    x = x.cuda()
    condition = torch.randn(1, 20, 2, 90, 160, generator=torch.Generator().manual_seed(5)).bfloat16()
    model_input = torch.empty(1, 36, 2, 90, 160, dtype=torch.bfloat16, device="cuda")
    model_input[:, 16:] = condition.cuda()
    for i, t in enumerate(s.timesteps):
        x = s.step_cfg(cond[i].cuda(), uncond[i].cuda(), case.guidance, t, x, model_input_out=model_input)
        assert torch.equal(_bits(x.cpu()), _bits(ora[i]))
        assert torch.equal(_bits(model_input.cpu()), _bits(unipc_oracle.model_input(ora[i], condition)))


def test_state_shape_mismatch_is_an_error():
    from chronoedit_b200.scheduler import FlowUniPCMultistepScheduler
    s = FlowUniPCMultistepScheduler(shift=1)
    s.set_timesteps(4, device="cuda", shift=2.0)
    x = torch.randn(1, 16, 4, 4, 6, device="cuda").bfloat16()
    x = s.step(torch.randn_like(x), s.timesteps[0], x, return_dict=False)[0]
    with pytest.raises(ValueError):   # the cut without slicing the scheduler state
        s.step(torch.randn_like(x[:, :, [0, -1]]), s.timesteps[1], x[:, :, [0, -1]])
    with pytest.raises(TypeError):
        s.step(torch.randn_like(x).float(), s.timesteps[1], x)
