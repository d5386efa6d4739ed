"""Parity of the two encoder mirrors against the classes the reference pipeline itself instantiates --
`transformers.UMT5EncoderModel` and `transformers.CLIPVisionModel` (pipeline_chronoedit.py:23, 205-254) -- run on the same
device with the same random weights: bf16 (the reference's configuration) and fp32 (the exact answer).  Same acceptance rule as the
DiT / VAE: |ours - fp32| <= 1.25 x |reference bf16 - fp32| in the mean, 2x in the max.  transformers is importable on the GPU box,
so here the reference is the real thing, not a restatement."""
import json
import os

import pytest
import torch

gpu = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(name, ours, ref16, ref32):
    e_ref, e_our = (ref16 - ref32).abs(), (ours - ref32).abs()
    rep = {"case": name, "ours_mean": e_our.mean().item(), "ref_bf16_mean": e_ref.mean().item(), "ours_max": e_our.max().item(),
           "ref_bf16_max": e_ref.max().item(), "mean_abs_ref32": ref32.abs().mean().item()}
    print(json.dumps(rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"encoder_parity_{name}.json"), "w") as f:
        json.dump(rep, f)
    assert torch.isfinite(ours).all()
    assert e_our.mean() <= 1.25 * e_ref.mean(), rep
    assert e_our.max() <= 2.0 * e_ref.max(), rep


@gpu
@pytest.mark.parametrize("name,kw,valid", [
    ("umt5_small", dict(vocab_size=4096, d_model=512, d_kv=64, d_ff=1024, num_layers=3, num_heads=8), [512, 37]),
    # google/umt5-xxl width (d_model 4096, 64 heads x 64, d_ff 10240), two layers, reduced vocabulary
    ("umt5_xxl_width", dict(vocab_size=8192, d_model=4096, d_kv=64, d_ff=10240, num_layers=2, num_heads=64), [200]),
])
def test_umt5_encoder_matches_transformers(name, kw, valid):
    from transformers import UMT5Config, UMT5EncoderModel

    import chronoedit_b200 as ce

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cfg = UMT5Config(feed_forward_proj="gated-gelu", relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0, **kw)
    ref = UMT5EncoderModel(cfg).eval()
    with torch.no_grad():
        for n, p in ref.named_parameters():   # transformers' init leaves tiny / huge scales; use O(1) activations
            if "layer_norm" in n:
                p.normal_(1.0, 0.1)
            elif "relative_attention_bias" in n:
                p.normal_(0.0, 1.0)
            elif "shared" in n or "embed_tokens" in n:
                p.normal_(0.0, 1.0)
            else:
                p.normal_(0.0, (1.0 / p.shape[1]) ** 0.5)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    sd16 = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    B, L = len(valid), 512
    ids = torch.randint(0, kw["vocab_size"], (B, L))
    mask = torch.zeros(B, L, dtype=torch.long)
    for b, v in enumerate(valid):
        mask[b, :v] = 1
    m = ce.UMT5EncoderModel(**kw)
    m.load_state_dict(sd16)
    m = m.cuda()
    ours = m(ids.cuda(), mask.cuda()).last_hidden_state.float()
    assert m.launches() > 0
    ref = ref.cuda()
    ref.load_state_dict({k: v.float() for k, v in sd16.items()})   # the same bf16-representable weights, computed exactly
    with torch.no_grad():
        r32 = ref(ids.cuda(), mask.cuda()).last_hidden_state.float()
        r16 = ref.to(torch.bfloat16)(ids.cuda(), mask.cuda()).last_hidden_state.float()
    for b, v in enumerate(valid):   # the pipeline keeps only the valid prefix (pipeline_chronoedit.py:234-237); padded positions are undefined
        _report(f"{name}_b{b}", ours[b, :v], r16[b, :v], r32[b, :v])


@gpu
@pytest.mark.parametrize("name,kw", [
    ("clip_small", dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4, image_size=56, patch_size=14, hidden_act="quick_gelu")),
    # ViT-H/14 as the reference loads it: 1280 wide, 16 heads x 80, 32 layers, 224 px -> 257 tokens
    ("clip_vit_h", dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224, patch_size=14, hidden_act="gelu")),
])
def test_clip_vision_encoder_matches_transformers(name, kw):
    from transformers import CLIPVisionConfig, CLIPVisionModel

    import chronoedit_b200 as ce

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    cfg = CLIPVisionConfig(attention_dropout=0.0, **kw)
    ref = CLIPVisionModel(cfg).eval()
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.normal_(1.0, 0.1)
            elif n.endswith("bias"):
                p.normal_(0.0, 0.05)
            elif "embedding" in n:
                p.normal_(0.0, 0.5)
            elif p.dim() == 2:
                p.normal_(0.0, (1.0 / p.shape[1]) ** 0.5)
    sd16 = {k: v.detach().to(torch.bfloat16) for k, v in ref.state_dict().items()}
    px = torch.randn(2, 3, kw["image_size"], kw["image_size"])
    m = ce.CLIPVisionModel(**kw)
    m.load_state_dict(sd16)
    m = m.cuda()
    ours = m(pixel_values=px.cuda(), output_hidden_states=True).hidden_states[-2].float()
    assert m.launches() > 0
    ref = ref.cuda()
    ref.load_state_dict({k: v.float() for k, v in sd16.items()})
    with torch.no_grad():
        r32 = ref(pixel_values=px.cuda(), output_hidden_states=True).hidden_states[-2].float()
        r16 = ref.to(torch.bfloat16)(pixel_values=px.cuda().bfloat16(), output_hidden_states=True).hidden_states[-2].float()
    assert ours.shape == r32.shape
    _report(name, ours, r16, r32)
