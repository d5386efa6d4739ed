#!/usr/bin/env python
"""Golden fixtures for the per-step sampling glue, recorded from the UNMODIFIED reference scheduler in this container.

    python tests/golden/make_golden_unipc.py      # writes tests/golden/unipc_<case>.safetensors + UNIPC_MANIFEST.json

Source of truth executed here (never copied): /root/reference/chronoedit/_src/models/fm_solvers_unipc.py
(`FlowUniPCMultistepScheduler`), with its un-vendored diffusers imports resolved by oracle/diffusers_shim.  The model is
replaced by seeded noise of the right shape/dtype: the scheduler is elementwise in the model output, so any values exercise
it.  Each case stores the initial sample, every step's model output (or cond/uncond pair) and every step's returned sample;
the temporal-reasoning case applies the slicing of chronoedit_diffusers/pipeline_chronoedit.py:700-709 to the reference
object's state exactly as the pipeline does.  The oracle restatement must reproduce every step bit for bit or generation
aborts."""
from __future__ import annotations

import json
import os
import sys

import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, unipc_oracle  # noqa: E402
from oracle.unipc_cases import UNIPC_CASES, UniPCCase, case_inputs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_reference(case: UniPCCase):
    ref = ref_loader.load_reference_unipc()
    sch = ref.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sch.set_timesteps(case.steps, device="cpu", shift=case.shift)
    x, cond, uncond = case_inputs(case)
    outs = []
    for i, t in enumerate(sch.timesteps):
        if case.cut_at is not None and i == case.cut_at:
            x = x[:, :, [0, -1]]
            for j in range(len(sch.model_outputs)):
                if sch.model_outputs[j] is not None and x.shape[-3] != sch.model_outputs[j].shape[-3]:
                    sch.model_outputs[j] = sch.model_outputs[j][:, :, [0, -1]]
            if sch.last_sample is not None:
                sch.last_sample = sch.last_sample[:, :, [0, -1]]
        c, u = cond[i], uncond[i]
        if case.cut_at is not None and i >= case.cut_at:
            c, u = c[:, :, [0, -1]], u[:, :, [0, -1]]
        v = u + case.guidance * (c - u) if case.guidance is not None else c   # pipeline_chronoedit.py:736
        x = sch.step(v, t, x, return_dict=False)[0]
        outs.append(x)
    return sch, outs


def run_oracle(case: UniPCCase, cuda_semantics=False):
    o = unipc_oracle.UniPCOracle(shift=1.0, cuda_semantics=cuda_semantics)
    o.set_timesteps(case.steps, shift=case.shift)
    x, cond, uncond = case_inputs(case)
    outs = []
    for i in range(case.steps):
        if case.cut_at is not None and i == case.cut_at:
            x = x[:, :, [0, -1]]
            o.cut_frames()
        c, u = cond[i], uncond[i]
        if case.cut_at is not None and i >= case.cut_at:
            c, u = c[:, :, [0, -1]], u[:, :, [0, -1]]
        v = unipc_oracle.cfg_combine(c, u, case.guidance, cuda_semantics) if case.guidance is not None else c
        x = o.step(v, x)
        outs.append(x)
    return o, outs


def main():
    manifest = {}
    for name, case in UNIPC_CASES.items():
        sch, ref_outs = run_reference(case)
        o, ora_outs = run_oracle(case)
        assert torch.equal(sch.timesteps, o.timesteps) and torch.equal(sch.sigmas, o.sigmas), name
        for i, (a, b) in enumerate(zip(ref_outs, ora_outs)):
            assert a.dtype == b.dtype and a.shape == b.shape, (name, i, a.dtype, b.dtype)
            assert torch.equal(a.view(torch.int16 if a.dtype == torch.bfloat16 else torch.int32),
                               b.view(torch.int16 if b.dtype == torch.bfloat16 else torch.int32)), f"{name}: oracle != reference at step {i}"
        tensors = {f"step{i:02d}": t.contiguous() for i, t in enumerate(ref_outs)}
        tensors["sigmas"] = sch.sigmas.clone()
        tensors["timesteps"] = sch.timesteps.clone()
        save_file(tensors, os.path.join(OUT, f"unipc_{name}.safetensors"))
        # how far torch's CUDA scalar semantics (fp32 coefficients, reciprocal multiply; see oracle/unipc_oracle.py) move the
        # result away from this CPU run of the reference: recorded for the GPU tests' tolerance, not asserted
        _, ora_cuda = run_oracle(case, cuda_semantics=True)
        n_diff = sum(int((a != b).sum()) for a, b in zip(ref_outs, ora_cuda))
        max_diff = max(float((a.float() - b.float()).abs().max()) for a, b in zip(ref_outs, ora_cuda))
        manifest[name] = {"steps": case.steps, "shift": case.shift, "shape": list(case.shape), "sample_dtype": str(case.sample_dtype),
                          "model_dtype": str(case.model_dtype), "guidance": case.guidance, "cut_at": case.cut_at,
                          "final_abs_mean": float(ref_outs[-1].float().abs().mean()),
                          "elements_changed_by_cuda_semantics": n_diff, "max_abs_change_by_cuda_semantics": max_diff,
                          "elements_total": int(sum(t.numel() for t in ref_outs))}
        print(name, "ok", manifest[name])
    with open(os.path.join(OUT, "UNIPC_MANIFEST.json"), "w") as f:
        json.dump({"source": "chronoedit/_src/models/fm_solvers_unipc.py (unmodified, CPU)", "torch": torch.__version__, "cases": manifest}, f, indent=1)


if __name__ == "__main__":
    main()
