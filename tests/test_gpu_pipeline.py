"""Pipeline-level parity on the GPU: the sampling loop of `ChronoEditPipeline.__call__` (restated in oracle/pipeline_oracle.py and
pinned bit for bit against the UNMODIFIED pipeline file in tests/test_pipeline_cpu.py -- the reference file itself does not
exist on the GPU box) drives

  (a) the three chronoedit_b200 mirrors (DiT, VAE, flow-UniPC scheduler): the CUDA path under test, through the C ABI;
  (b) the oracle modules evaluated with torch on the same device in the reference's bf16 configuration (= what the
      unmodified reference computes on a GPU: cuBLAS / cuDNN / SDPA and torch's CUDA scalar semantics in the scheduler);
  (c) the oracle modules in fp32 (TF32 off): the exact answer.

Acceptance = tests/test_gpu_dit.py: |(a) - (c)| <= 1.25 x |(b) - (c)| in the mean (2x in the max) on the decoded video, for
a plain 5-frame edit, the guidance-free (distilled-LoRA style) schedule, and temporal reasoning with the in-loop cut
(pipeline_chronoedit.py:700-709) and with the CLI default (no cut, two decodes, :776-779)."""
import json
import os

import pytest
import torch

gpu = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mirrors(case, dev):
    import chronoedit_b200 as ce
    from chronoedit_b200.autoencoder import AutoencoderKLWan
    from oracle import cases, pipeline_cases as PC

    dsd, vsd = PC.weights()
    cfg = PC.DIT_CFG
    tr = ce.ChronoEditTransformer3DModel(
        patch_size=cfg.patch_size, num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, text_dim=cfg.text_dim, freq_dim=cfg.freq_dim, ffn_dim=cfg.ffn_dim,
        num_layers=cfg.num_layers, eps=cfg.eps, image_dim=cfg.image_dim, added_kv_proj_dim=cfg.added_kv_proj_dim, cache_context=True)
    tr.load_state_dict(cases.to_bf16_state(dsd))
    vc = PC.VAE_CFG
    vae = AutoencoderKLWan(base_dim=vc.dim, z_dim=vc.z_dim, dim_mult=tuple(vc.dim_mult), num_res_blocks=vc.num_res_blocks,
                           temperal_downsample=tuple(vc.temperal_downsample))
    vae.load_state_dict({k: v.to(torch.bfloat16) for k, v in vsd.items()})
    return tr.to(dev), vae.to(dev), ce.FlowUniPCMultistepScheduler(shift=case.sched_shift)


@gpu
@pytest.mark.parametrize("name", ["edit_5f", "edit_nocfg", "reason_cut", "reason_full"])
def test_pipeline_with_the_cuda_mirrors(name):
    from oracle import pipeline_cases as PC

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    case = PC.PIPELINE_CASES[name]
    tr, vae, sch = _mirrors(case, dev)
    ours = PC.run_oracle_pipeline(case, torch.bfloat16, device=dev, transformer=tr, vae=vae, scheduler=sch).float()
    assert tr.launches_per_forward() > 0 and vae.launches() > 0
    ref16 = PC.run_oracle_pipeline(case, torch.bfloat16, device=dev).float()
    ref32 = PC.run_oracle_pipeline(case, torch.float32, device=dev).float()
    torch.cuda.synchronize()
    assert ours.shape == ref32.shape and torch.isfinite(ours).all()
    e_ref, e_our = (ref16 - ref32).abs(), (ours - ref32).abs()
    rep = {"case": name, "ours_mean": e_our.mean().item(), "ref_bf16_mean": e_ref.mean().item(), "ours_max": e_our.max().item(),
           "ref_bf16_max": e_ref.max().item(), "video_mean_abs": ref32.abs().mean().item()}
    print(json.dumps(rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"pipeline_parity_{name}.json"), "w") as f:
        json.dump(rep, f)
    assert e_our.mean() <= 1.25 * e_ref.mean(), f"mean err {e_our.mean():.3g} vs the reference's own bf16 error {e_ref.mean():.3g}"
    assert e_our.max() <= 2.0 * e_ref.max() + 1e-3, f"max err {e_our.max():.3g} vs the reference's own bf16 error {e_ref.max():.3g}"
