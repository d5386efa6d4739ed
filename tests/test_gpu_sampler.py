"""GPU parity tests (-m gpu) of the fused sampling-glue launch (csrc/sampler.cu through ce_unipc_step and the
`FlowUniPCMultistepScheduler` mirror): bit-exact against the oracle (CUDA division semantics) on every step of every
case, bit-exact against the golden vectors of the unmodified reference where its CPU division cannot show (all bf16 cases),
and at the full 720p latent size."""
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


def _run_oracle(case, recip_div=True):
    o = unipc_oracle.UniPCOracle(shift=1.0, recip_div=recip_div)
    o.set_timesteps(case.steps, shift=case.shift)
    x, cond, uncond = case_inputs(case)
    outs, x0s = [], []
    for i in range(case.steps):
        if case.cut_at is not None and i == case.cut_at:
            x = x[:, :, [0, -1]]
            o.cut_frames()
        c, u = cond[i], uncond[i]
        if case.cut_at is not None and i >= case.cut_at:
            c, u = c[:, :, [0, -1]], u[:, :, [0, -1]]
        v = unipc_oracle.cfg_combine(c, u, case.guidance) if case.guidance is not None else c
        x = o.step(v, x)
        outs.append(x)
        x0s.append(o.model_outputs[-1])
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
    _, ora, ora_x0 = _run_oracle(case, recip_div=True)
    s, got, got_x0 = _run_mirror(case, fused_cfg)
    gold = load_file(os.path.join(golden_dir, f"unipc_{name}.safetensors"))
    man = json.load(open(os.path.join(golden_dir, "UNIPC_MANIFEST.json")))["cases"][name]
    for i in range(case.steps):
        assert got[i].dtype == ora[i].dtype and got[i].shape == ora[i].shape
        assert torch.equal(_bits(got[i].cpu()), _bits(ora[i])), f"{name}: sample after step {i} differs from the oracle"
        assert torch.equal(_bits(got_x0[i].cpu()), _bits(ora_x0[i])), f"{name}: x0 prediction of step {i} differs from the oracle"
        g = gold[f"step{i:02d}"]
        if man["elements_changed_by_reciprocal_division"] == 0:
            assert torch.equal(_bits(got[i].cpu()), _bits(g)), f"{name}: step {i} differs from the unmodified reference"
        else:   # fp32: the reference's CPU run divides where its CUDA run (and this kernel) multiplies by 1/r_k
            torch.testing.assert_close(got[i].cpu(), g, rtol=2e-6, atol=2e-6)
    assert s.step_index == case.steps and s.lower_order_nums == min(case.steps, 2)


def test_full_size_cfg_step_with_model_input():
    """720p / 5-frame latent [1,16,2,90,160], bf16, CFG 5.0, 6 steps; the fused launch also writes the latent channels of the
    next model input (pipeline_chronoedit.py:712)."""
    from chronoedit_b200.scheduler import FlowUniPCMultistepScheduler
    case = UniPCCase(6, 5.0, (1, 16, 2, 90, 160), torch.bfloat16, torch.bfloat16, guidance=5.0, seed=21)
    _, ora, _ = _run_oracle(case, recip_div=True)
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
