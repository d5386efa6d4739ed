"""CPU restatement of the ChronoEdit DiT per-step forward.

TEST INFRASTRUCTURE ONLY — the checker, never the product (see oracle/__init__.py).

Restates, as plain functional PyTorch on a state-dict, the algorithm of
  /root/reference/chronoedit_diffusers/transformer_chronoedit.py
(ChronoEditTransformer3DModel.forward, :397-476) including every dtype/rounding
point of the reference's bf16 path.  Parameter names are the reference module
tree's (`blocks.N.attn1.to_q.weight`, ...), so a reference state_dict can be fed
in unchanged.

Pinning: `tests/golden/make_golden.py` runs the UNMODIFIED reference file (through
oracle/diffusers_shim) and this restatement on the same seeded weights/inputs and
requires agreement (fp32: <=1e-5 abs; bf16: bit-exact, same CPU kernels), then
stores the reference outputs as fixtures under tests/golden/.  The reference's
own tests hold no vectors for this path, and diffusers
itself is un-vendored, so the diffusers pieces are additionally cross-checked
against the in-tree DiffSynth implementation (wan_video_dit_chronoedit.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class DiTConfig:
    """Hyper-parameters (transformer_chronoedit.py:342-360)."""

    patch_size: Tuple[int, int, int] = (1, 2, 2)
    num_attention_heads: int = 40
    attention_head_dim: int = 128
    in_channels: int = 36
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 13824
    num_layers: int = 40
    cross_attn_norm: bool = True
    qk_norm: Optional[str] = "rms_norm_across_heads"
    eps: float = 1e-6
    image_dim: Optional[int] = 1280
    added_kv_proj_dim: Optional[int] = 5120
    rope_max_seq_len: int = 1024
    rope_temporal_skip_len: int = 8

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @staticmethod
    def chronoedit_14b() -> "DiTConfig":
        return DiTConfig()

    @staticmethod
    def tiny(num_layers: int = 2, heads: int = 2, ffn_dim: int = 1024, text_dim: int = 4096,
             image_dim: int = 1280, rope_temporal_skip_len: int = 8) -> "DiTConfig":
        """BASELINE.json configs[0]: 2-layer / dim-256 (2 heads x 128)."""
        d = heads * 128
        return DiTConfig(num_attention_heads=heads, attention_head_dim=128, ffn_dim=ffn_dim, num_layers=num_layers,
                         text_dim=text_dim, image_dim=image_dim, added_kv_proj_dim=d,
                         rope_temporal_skip_len=rope_temporal_skip_len)


# names kept in fp32 when the model is cast to bf16 (transformer_chronoedit.py:338)
KEEP_FP32 = ("time_embedder", "scale_shift_table", "norm1", "norm2", "norm3")


def param_shapes(cfg: DiTConfig) -> Dict[str, Tuple[int, ...]]:
    """Every parameter of the reference module tree, name -> shape (order = reference registration order)."""
    D, F_, H = cfg.inner_dim, cfg.ffn_dim, cfg.freq_dim
    pt, ph, pw = cfg.patch_size
    s: Dict[str, Tuple[int, ...]] = {}
    s["patch_embedding.weight"] = (D, cfg.in_channels, pt, ph, pw)
    s["patch_embedding.bias"] = (D,)
    ce = "condition_embedder."
    s[ce + "time_embedder.linear_1.weight"] = (D, H)
    s[ce + "time_embedder.linear_1.bias"] = (D,)
    s[ce + "time_embedder.linear_2.weight"] = (D, D)
    s[ce + "time_embedder.linear_2.bias"] = (D,)
    s[ce + "time_proj.weight"] = (6 * D, D)
    s[ce + "time_proj.bias"] = (6 * D,)
    s[ce + "text_embedder.linear_1.weight"] = (D, cfg.text_dim)
    s[ce + "text_embedder.linear_1.bias"] = (D,)
    s[ce + "text_embedder.linear_2.weight"] = (D, D)
    s[ce + "text_embedder.linear_2.bias"] = (D,)
    if cfg.image_dim is not None:
        I = cfg.image_dim
        s[ce + "image_embedder.norm1.weight"] = (I,)
        s[ce + "image_embedder.norm1.bias"] = (I,)
        s[ce + "image_embedder.ff.net.0.proj.weight"] = (I, I)
        s[ce + "image_embedder.ff.net.0.proj.bias"] = (I,)
        s[ce + "image_embedder.ff.net.2.weight"] = (D, I)
        s[ce + "image_embedder.ff.net.2.bias"] = (D,)
        s[ce + "image_embedder.norm2.weight"] = (D,)
        s[ce + "image_embedder.norm2.bias"] = (D,)
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        for a in ("attn1", "attn2"):
            s[b + a + ".norm_q.weight"] = (D,)
            s[b + a + ".norm_k.weight"] = (D,)
            for p in ("to_q", "to_k", "to_v"):
                s[b + a + f".{p}.weight"] = (D, D)
                s[b + a + f".{p}.bias"] = (D,)
            if a == "attn2" and cfg.added_kv_proj_dim is not None:
                s[b + a + ".add_k_proj.weight"] = (D, cfg.added_kv_proj_dim)
                s[b + a + ".add_k_proj.bias"] = (D,)
                s[b + a + ".add_v_proj.weight"] = (D, cfg.added_kv_proj_dim)
                s[b + a + ".add_v_proj.bias"] = (D,)
                s[b + a + ".norm_added_k.weight"] = (D,)
            s[b + a + ".to_out.0.weight"] = (D, D)
            s[b + a + ".to_out.0.bias"] = (D,)
        if cfg.cross_attn_norm:
            s[b + "norm2.weight"] = (D,)
            s[b + "norm2.bias"] = (D,)
        s[b + "ffn.net.0.proj.weight"] = (F_, D)
        s[b + "ffn.net.0.proj.bias"] = (F_,)
        s[b + "ffn.net.2.weight"] = (D, F_)
        s[b + "ffn.net.2.bias"] = (D,)
        s[b + "scale_shift_table"] = (1, 6, D)
    s["proj_out.weight"] = (cfg.out_channels * pt * ph * pw, D)
    s["proj_out.bias"] = (cfg.out_channels * pt * ph * pw,)
    s["scale_shift_table"] = (1, 2, D)
    return s


def random_state_dict(cfg: DiTConfig, seed: int = 0, dtype: torch.dtype = torch.float32) -> Dict[str, Tensor]:
    """Seeded synthetic weights (SURVEY.md section 8c "How we pin it").

    Linear/conv weights ~ N(0, 0.02), biases ~ N(0, 0.02), scale_shift_table ~ N(0,1)/sqrt(D)
    (as transformer_chronoedit.py:265,393), norm weights 1 + 0.1 N(0,1) so the affine paths matter.
    Generated in fp32 on the CPU generator, then cast; KEEP_FP32 names stay fp32.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("scale_shift_table"):
            w = torch.randn(shape, generator=g) / math.sqrt(cfg.inner_dim)
        elif ".norm" in name and name.endswith(".weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".norm" in name and name.endswith(".bias"):
            w = 0.1 * torch.randn(shape, generator=g)
        else:
            w = 0.02 * torch.randn(shape, generator=g)
        keep = any(k in name for k in KEEP_FP32)
        sd[name] = w if (keep or dtype == torch.float32) else w.to(dtype)
    return sd


# ----------------------------------------------------------------------------------------------
# pieces
# ----------------------------------------------------------------------------------------------

def rope_table(cfg: DiTConfig, frames: int, height: int, width: int) -> Tensor:
    """complex128 [1,1,L,hd/2] rotary table — ChronoEditRotaryPosEmbed (transformer_chronoedit.py:168-213).

    The head dim splits into (t,h,w) = (hd - 4*(hd//6), 2*(hd//6), 2*(hd//6)) real dims; frequencies are
    1/theta^(2i/dim) in fp64 (diffusers get_1d_rotary_pos_embed).  With exactly 2 latent frames the temporal
    positions are {0, temporal_skip_len-1} (:206-207), otherwise 0..frames-1 and frames must equal
    temporal_skip_len (:205).
    """
    hd = cfg.attention_head_dim
    pt, ph, pw = cfg.patch_size
    ppf, pph, ppw = frames // pt, height // ph, width // pw
    if not (frames == 2 or frames == cfg.rope_temporal_skip_len):
        raise AssertionError(f"num_frames must be 2 or {cfg.rope_temporal_skip_len}, but got {frames}")
    h_dim = w_dim = 2 * (hd // 6)
    t_dim = hd - h_dim - w_dim

    def axis(dim: int) -> Tensor:
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
        ang = torch.outer(torch.arange(cfg.rope_max_seq_len), inv)
        return torch.polar(torch.ones_like(ang), ang)

    ft, fh, fw = axis(t_dim), axis(h_dim), axis(w_dim)
    if frames == 2:
        ft = ft[: cfg.rope_temporal_skip_len][[0, -1]]
    else:
        ft = ft[:ppf]
    ft = ft.view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1)
    fh = fh[:pph].view(1, pph, 1, -1).expand(ppf, pph, ppw, -1)
    fw = fw[:ppw].view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1)
    return torch.cat([ft, fh, fw], dim=-1).reshape(1, 1, ppf * pph * ppw, -1)


def _rms_norm_across_heads(x: Tensor, weight: Tensor, eps: float) -> Tensor:
    """diffusers RMSNorm over the full inner dim (transformer_chronoedit.py:62-65).
    fp32 variance; x*rstd is fp32, cast to the weight dtype when that is bf16/fp16, then * weight."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        y = y.to(weight.dtype)
    return y * weight


def _fp32_layer_norm(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float) -> Tensor:
    """diffusers FP32LayerNorm: statistics and affine in fp32, result cast back to the input dtype."""
    return F.layer_norm(x.float(), (x.shape[-1],), None if weight is None else weight.float(),
                        None if bias is None else bias.float(), eps).to(x.dtype)


def _apply_rope(x: Tensor, freqs: Tensor) -> Tensor:
    """x [B,H,L,hd]; pairs (2i,2i+1) rotated in complex128, cast back (transformer_chronoedit.py:73-79)."""
    xc = torch.view_as_complex(x.to(torch.float64).unflatten(3, (-1, 2)))
    return torch.view_as_real(xc * freqs).flatten(3, 4).type_as(x)


def _heads(x: Tensor, h: int) -> Tensor:
    return x.unflatten(2, (h, -1)).transpose(1, 2)


def attention(sd: Dict[str, Tensor], p: str, cfg: DiTConfig, x: Tensor, ctx: Optional[Tensor],
              freqs: Optional[Tensor]) -> Tensor:
    """ChronoEditAttnProcessor2_0.__call__ (transformer_chronoedit.py:43-108).
    `ctx` None => self-attention; otherwise ctx = [image tokens (257) ; text tokens]."""
    H = cfg.num_attention_heads
    has_img = (p + "add_k_proj.weight") in sd
    img = None
    if ctx is not None and has_img:
        img, ctx = ctx[:, :257], ctx[:, 257:]
    if ctx is None:
        ctx = x
    q = F.linear(x, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
    k = F.linear(ctx, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
    v = F.linear(ctx, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
    q = _rms_norm_across_heads(q, sd[p + "norm_q.weight"], cfg.eps)
    k = _rms_norm_across_heads(k, sd[p + "norm_k.weight"], cfg.eps)
    q, k, v = _heads(q, H), _heads(k, H), _heads(v, H)
    if freqs is not None:
        q = _apply_rope(q, freqs)
        k = _apply_rope(k, freqs)
    o_img = None
    if img is not None:
        ki = F.linear(img, sd[p + "add_k_proj.weight"], sd[p + "add_k_proj.bias"])
        ki = _rms_norm_across_heads(ki, sd[p + "norm_added_k.weight"], cfg.eps)
        vi = F.linear(img, sd[p + "add_v_proj.weight"], sd[p + "add_v_proj.bias"])
        o_img = F.scaled_dot_product_attention(q, _heads(ki, H), _heads(vi, H))
        o_img = o_img.transpose(1, 2).flatten(2, 3).type_as(q)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).flatten(2, 3).type_as(q)
    if o_img is not None:
        o = o + o_img
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def block(sd: Dict[str, Tensor], i: int, cfg: DiTConfig, x: Tensor, ctx: Tensor, temb6: Tensor,
          freqs: Tensor) -> Tensor:
    """ChronoEditTransformerBlock.forward (transformer_chronoedit.py:267-295)."""
    p = f"blocks.{i}."
    shift_msa, scale_msa, gate_msa, c_shift, c_scale, c_gate = (sd[p + "scale_shift_table"] + temb6.float()).chunk(
        6, dim=1)
    # 1. self-attention: fp32 LN + modulate -> model dtype (:279); gated residual in fp32 (:281)
    n = (_fp32_layer_norm(x.float(), None, None, cfg.eps) * (1 + scale_msa) + shift_msa).type_as(x)
    a = attention(sd, p + "attn1.", cfg, n, None, freqs)
    x = (x.float() + a * gate_msa).type_as(x)
    # 2. cross-attention: affine fp32 LN (:284), residual add in model dtype (:286)
    if cfg.cross_attn_norm:
        n = _fp32_layer_norm(x.float(), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.eps).type_as(x)
    else:
        n = x.float().type_as(x)
    a = attention(sd, p + "attn2.", cfg, n, ctx, None)
    x = x + a
    # 3. feed-forward: Linear -> GELU(tanh) -> Linear (:292), gated residual in fp32 (:293)
    n = (_fp32_layer_norm(x.float(), None, None, cfg.eps) * (1 + c_scale) + c_shift).type_as(x)
    f = F.linear(n, sd[p + "ffn.net.0.proj.weight"], sd[p + "ffn.net.0.proj.bias"])
    f = F.gelu(f, approximate="tanh")
    f = F.linear(f, sd[p + "ffn.net.2.weight"], sd[p + "ffn.net.2.bias"])
    x = (x.float() + f.float() * c_gate).type_as(x)
    return x


def timestep_sinusoid(timestep: Tensor, dim: int) -> Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], fp32."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timestep.device) / half
    emb = timestep[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def condition_embedder(sd: Dict[str, Tensor], cfg: DiTConfig, timestep: Tensor, text: Tensor,
                       image: Optional[Tensor]):
    """ChronoEditTimeTextImageEmbedding.forward (transformer_chronoedit.py:147-165) and
    ChronoEditImageEmbedding.forward (:111-123)."""
    p = "condition_embedder."
    t = timestep_sinusoid(timestep, cfg.freq_dim)
    w1 = sd[p + "time_embedder.linear_1.weight"]
    if t.dtype != w1.dtype:
        t = t.to(w1.dtype)
    temb = F.linear(F.silu(F.linear(t, w1, sd[p + "time_embedder.linear_1.bias"])),
                    sd[p + "time_embedder.linear_2.weight"], sd[p + "time_embedder.linear_2.bias"]).type_as(text)
    tproj = F.linear(F.silu(temb), sd[p + "time_proj.weight"], sd[p + "time_proj.bias"])
    text = F.linear(F.gelu(F.linear(text, sd[p + "text_embedder.linear_1.weight"],
                                    sd[p + "text_embedder.linear_1.bias"]), approximate="tanh"),
                    sd[p + "text_embedder.linear_2.weight"], sd[p + "text_embedder.linear_2.bias"])
    if image is not None:
        q = p + "image_embedder."
        image = _fp32_layer_norm(image, sd[q + "norm1.weight"], sd[q + "norm1.bias"], 1e-5)
        image = F.gelu(F.linear(image, sd[q + "ff.net.0.proj.weight"], sd[q + "ff.net.0.proj.bias"]))
        image = F.linear(image, sd[q + "ff.net.2.weight"], sd[q + "ff.net.2.bias"])
        image = _fp32_layer_norm(image, sd[q + "norm2.weight"], sd[q + "norm2.bias"], 1e-5)
    return temb, tproj, text, image


def dit_forward(sd: Dict[str, Tensor], cfg: DiTConfig, hidden_states: Tensor, timestep: Tensor,
                encoder_hidden_states: Tensor, encoder_hidden_states_image: Optional[Tensor] = None,
                num_layers: Optional[int] = None, return_intermediates: bool = False):
    """ChronoEditTransformer3DModel.forward (transformer_chronoedit.py:397-476).

    hidden_states [B,Cin,T,H,W]; timestep [B]; encoder_hidden_states [B,Lt,text_dim];
    encoder_hidden_states_image [B,257,image_dim] -> sample [B,Cout,T,H,W].
    """
    B, _, T, Hh, Ww = hidden_states.shape
    pt, ph, pw = cfg.patch_size
    ppf, pph, ppw = T // pt, Hh // ph, Ww // pw
    freqs = rope_table(cfg, T, Hh, Ww).to(hidden_states.device)  # the oracle also runs on CUDA tensors (tests at BASELINE sizes)
    x = F.conv3d(hidden_states, sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    temb, tproj, text, image = condition_embedder(sd, cfg, timestep, encoder_hidden_states,
                                                  encoder_hidden_states_image)
    temb6 = tproj.unflatten(1, (6, -1))
    ctx = text if image is None else torch.concat([image, text], dim=1)
    inter = {"patch_embed": x, "temb": temb, "timestep_proj": tproj, "context": ctx}
    n_layers = cfg.num_layers if num_layers is None else num_layers
    for i in range(n_layers):
        x = block(sd, i, cfg, x, ctx, temb6, freqs)
        if return_intermediates:
            inter[f"block{i}"] = x
    shift, scale = (sd["scale_shift_table"] + temb.unsqueeze(1)).chunk(2, dim=1)
    x = (_fp32_layer_norm(x.float(), None, None, cfg.eps) * (1 + scale) + shift).type_as(x)
    x = F.linear(x, sd["proj_out.weight"], sd["proj_out.bias"])
    x = x.reshape(B, ppf, pph, ppw, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
    out = x.flatten(6, 7).flatten(4, 5).flatten(2, 3)
    if return_intermediates:
        return out, inter
    return out


def flops_per_forward(cfg: DiTConfig, frames: int, height: int, width: int, text_len: int = 512,
                      image_len: int = 257, batch: int = 1) -> float:
    """Algorithmic FLOPs of one forward (2*MAC; attention = 4*Lq*Lk*D) — SURVEY.md section 8d."""
    pt, ph, pw = cfg.patch_size
    L = (frames // pt) * (height // ph) * (width // pw)
    D, Fd = cfg.inner_dim, cfg.ffn_dim
    per_block = (
        2 * L * D * 3 * D            # self q,k,v
        + 4 * L * L * D              # self SDPA
        + 2 * L * D * D              # self out
        + 2 * L * D * D              # cross q
        + 2 * text_len * D * 2 * D   # cross k,v text
        + 2 * image_len * D * 2 * D  # cross k,v image
        + 4 * L * (text_len + image_len) * D
        + 2 * L * D * D              # cross out
        + 2 * 2 * L * D * Fd         # ffn
    )
    embed = (2 * L * cfg.in_channels * pt * ph * pw * D + 2 * L * D * cfg.out_channels * pt * ph * pw
             + 2 * text_len * (cfg.text_dim * D + D * D)
             + 2 * image_len * ((cfg.image_dim or 0) ** 2 + (cfg.image_dim or 0) * D)
             + 2 * (cfg.freq_dim * D + D * D + D * 6 * D))
    return float(batch) * (cfg.num_layers * per_block + embed)
