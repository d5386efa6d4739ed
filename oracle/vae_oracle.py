"""CPU restatement of the Wan2.1 3D causal VAE encode/decode that bookend the sampling loop.

TEST INFRASTRUCTURE ONLY — the checker, never the product (see oracle/__init__.py).

The pipeline calls diffusers `AutoencoderKLWan` (un-vendored; call sites
/root/reference/chronoedit_diffusers/pipeline_chronoedit.py:436-443, 776-781).  Its arithmetic
twin lives in-tree at /root/reference/chronoedit/_src/tokenizers/wan2pt1.py and is what
this file restates (functional, explicit per-convolution stream state instead of the
reference's positional `feat_cache` list):

  CausalConv3d  wan2pt1.py:42-60     RMS_norm      :63-75      Upsample   :78-83
  Resample      :86-160              ResidualBlock :186-220    AttentionBlock :223-259
  Encoder3d     :262-357             Decoder3d     :360-456    WanVAE_.encode/.decode :502-560

Deltas of the diffusers class w.r.t. the in-tree twin that the drop-in boundary needs
(SURVEY.md section 8c): encode/decode apply NO latent mean/std (the pipeline does,
pipeline_chronoedit.py:427-445, 765-774), `encode` returns the 16-channel mean (mode of the
posterior), `decode` clamps to [-1, 1]  ([diffusers-mem], flag `clamp=`).

Parameter names are the in-tree twin's (`encoder.downsamples.3.residual.2.weight`, ...), so
the reference module's state_dict can be fed in unchanged.  Pinned by
tests/golden/make_golden.py against the UNMODIFIED reference module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
CACHE_T = 2  # wan2pt1.py:38


@dataclass
class VAEConfig:
    """_video_vae cfg (wan2pt1.py:597-605)."""

    dim: int = 96
    z_dim: int = 16
    dim_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temperal_downsample: Tuple[bool, ...] = (False, True, True)
    temporal_window: int = 4

    @staticmethod
    def wan21() -> "VAEConfig":
        return VAEConfig()

    @staticmethod
    def tiny(dim: int = 32) -> "VAEConfig":
        return VAEConfig(dim=dim)


# latent statistics used by the pipeline (values: wan2pt1.py:697-732)
LATENTS_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
                0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
LATENTS_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
               3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


# ----------------------------------------------------------------------------------------------
# architecture description (shared by encoder / decoder walkers and by param_shapes)
# ----------------------------------------------------------------------------------------------

def encoder_layers(cfg: VAEConfig) -> List[tuple]:
    """Encoder3d.__init__ (wan2pt1.py:281-313): list of ('res', name, cin, cout) / ('down', name, c, mode)."""
    dims = [cfg.dim * u for u in (1,) + tuple(cfg.dim_mult)]
    layers, idx = [], 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg.num_res_blocks):
            layers.append(("res", f"encoder.downsamples.{idx}", cin, cout))
            idx += 1
            cin = cout
        if i != len(cfg.dim_mult) - 1:
            mode = "downsample3d" if cfg.temperal_downsample[i] else "downsample2d"
            layers.append(("resample", f"encoder.downsamples.{idx}", cout, mode))
            idx += 1
    return layers


def decoder_layers(cfg: VAEConfig) -> List[tuple]:
    """Decoder3d.__init__ (wan2pt1.py:379-408)."""
    dims = [cfg.dim * u for u in (cfg.dim_mult[-1],) + tuple(cfg.dim_mult[::-1])]
    up = tuple(cfg.temperal_downsample[::-1])
    layers, idx = [], 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(cfg.num_res_blocks + 1):
            layers.append(("res", f"decoder.upsamples.{idx}", cin, cout))
            idx += 1
            cin = cout
        if i != len(cfg.dim_mult) - 1:
            mode = "upsample3d" if up[i] else "upsample2d"
            layers.append(("resample", f"decoder.upsamples.{idx}", cout, mode))
            idx += 1
    return layers


def param_shapes(cfg: VAEConfig) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}

    def conv3(name, cout, cin, k):
        s[name + ".weight"] = (cout, cin) + tuple(k)
        s[name + ".bias"] = (cout,)

    def res(name, cin, cout):
        s[name + ".residual.0.gamma"] = (cin, 1, 1, 1)
        conv3(name + ".residual.2", cout, cin, (3, 3, 3))
        s[name + ".residual.3.gamma"] = (cout, 1, 1, 1)
        conv3(name + ".residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv3(name + ".shortcut", cout, cin, (1, 1, 1))

    def attn(name, c):
        s[name + ".norm.gamma"] = (c, 1, 1)
        s[name + ".to_qkv.weight"] = (3 * c, c, 1, 1)
        s[name + ".to_qkv.bias"] = (3 * c,)
        s[name + ".proj.weight"] = (c, c, 1, 1)
        s[name + ".proj.bias"] = (c,)

    def resample(name, c, mode):
        if mode.startswith("upsample"):
            s[name + ".resample.1.weight"] = (c // 2, c, 3, 3)
            s[name + ".resample.1.bias"] = (c // 2,)
            if mode == "upsample3d":
                conv3(name + ".time_conv", 2 * c, c, (3, 1, 1))
        else:
            s[name + ".resample.1.weight"] = (c, c, 3, 3)
            s[name + ".resample.1.bias"] = (c,)
            if mode == "downsample3d":
                conv3(name + ".time_conv", c, c, (3, 1, 1))

    top = cfg.dim * cfg.dim_mult[-1]
    conv3("encoder.conv1", cfg.dim, 3, (3, 3, 3))
    for l in encoder_layers(cfg):
        res(l[1], l[2], l[3]) if l[0] == "res" else resample(l[1], l[2], l[3])
    res("encoder.middle.0", top, top)
    attn("encoder.middle.1", top)
    res("encoder.middle.2", top, top)
    s["encoder.head.0.gamma"] = (top, 1, 1, 1)
    conv3("encoder.head.2", 2 * cfg.z_dim, top, (3, 3, 3))
    conv3("conv1", 2 * cfg.z_dim, 2 * cfg.z_dim, (1, 1, 1))
    conv3("conv2", cfg.z_dim, cfg.z_dim, (1, 1, 1))
    conv3("decoder.conv1", top, cfg.z_dim, (3, 3, 3))
    res("decoder.middle.0", top, top)
    attn("decoder.middle.1", top)
    res("decoder.middle.2", top, top)
    for l in decoder_layers(cfg):
        res(l[1], l[2], l[3]) if l[0] == "res" else resample(l[1], l[2], l[3])
    s["decoder.head.0.gamma"] = (cfg.dim, 1, 1, 1)
    conv3("decoder.head.2", 3, cfg.dim, (3, 3, 3))
    return s


def random_state_dict(cfg: VAEConfig, seed: int = 0, dtype: torch.dtype = torch.float32) -> Dict[str, Tensor]:
    """Seeded synthetic weights: conv weights ~ N(0, 1/sqrt(fan_in)) so activations keep O(1) scale through
    the ~30-conv stack, biases ~ N(0, 0.02), gammas 1 + 0.1 N(0,1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("gamma"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            w = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[name] = w.to(dtype)
    return sd


# ----------------------------------------------------------------------------------------------
# streaming primitives
# ----------------------------------------------------------------------------------------------

class _Streams:
    """Per-convolution history of a causal stream: the last <=2 input frames seen so far
    (the reference's feat_cache entries, wan2pt1.py:204-217), keyed by layer name."""

    def __init__(self):
        self.state: Dict[str, object] = {}


def _causal_conv3d(x: Tensor, w: Tensor, b: Tensor, hist: Optional[Tensor], stride=(1, 1, 1)) -> Tensor:
    """CausalConv3d.forward (wan2pt1.py:53-60): symmetric spatial pad k//2, (k_t - 1) frames on the PAST side
    only (= 2*padding[0] for every conv the VAE builds); cached frames replace that many zero frames."""
    kt, kh, kw = w.shape[2:]
    pad_t, ph, pw = kt - 1, kh // 2, kw // 2
    if hist is not None and pad_t > 0:
        x = torch.cat([hist.to(x.device), x], dim=2)
        pad_t -= hist.shape[2]
    x = F.pad(x, (pw, pw, ph, ph, pad_t, 0))
    return F.conv3d(x, w, b, stride=stride)


def _push_history(x: Tensor, prev: Optional[Tensor]) -> Tensor:
    """New history after consuming chunk x: last CACHE_T frames; a 1-frame chunk keeps the last frame of
    the previous history in front of it (wan2pt1.py:206-212)."""
    h = x[:, :, -CACHE_T:].clone()
    if h.shape[2] < 2 and prev is not None:
        h = torch.cat([prev[:, :, -1:].to(h.device), h], dim=2)
    return h


def _stream_conv(st: _Streams, key: str, sd, name: str, x: Tensor) -> Tensor:
    prev = st.state.get(key)
    y = _causal_conv3d(x, sd[name + ".weight"], sd[name + ".bias"], prev)
    st.state[key] = _push_history(x, prev)
    return y


def _rms_norm(x: Tensor, gamma: Tensor) -> Tensor:
    """RMS_norm.forward (wan2pt1.py:74-75): L2-normalise over channels, * sqrt(C) * gamma."""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def _res_block(st: _Streams, sd, name: str, cin: int, cout: int, x: Tensor) -> Tensor:
    """ResidualBlock.forward (wan2pt1.py:201-220)."""
    h = x if cin == cout else _causal_conv3d(x, sd[name + ".shortcut.weight"], sd[name + ".shortcut.bias"], None)
    y = F.silu(_rms_norm(x, sd[name + ".residual.0.gamma"]))
    y = _stream_conv(st, name + ".residual.2", sd, name + ".residual.2", y)
    y = F.silu(_rms_norm(y, sd[name + ".residual.3.gamma"]))
    y = _stream_conv(st, name + ".residual.6", sd, name + ".residual.6", y)
    return y + h


def _attn_block(sd, name: str, x: Tensor) -> Tensor:
    """AttentionBlock.forward (wan2pt1.py:240-259): per-frame single-head attention over h*w tokens."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.normalize(y, dim=1) * (c ** 0.5) * sd[name + ".norm.gamma"]
    qkv = F.conv2d(y, sd[name + ".to_qkv.weight"], sd[name + ".to_qkv.bias"])
    q, k, v = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
    o = F.conv2d(o, sd[name + ".proj.weight"], sd[name + ".proj.bias"])
    o = o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
    return o + x


def _per_frame(x: Tensor, fn) -> Tensor:
    b, c, t, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def _resample(st: _Streams, sd, name: str, c: int, mode: str, x: Tensor) -> Tensor:
    """Resample.forward (wan2pt1.py:112-160)."""
    b, _, t, h, w = x.shape
    if mode == "upsample3d":
        key = name + ".time_conv"
        if key not in st.state:
            st.state[key] = "Rep"            # first chunk: no temporal doubling (:116-120)
        else:
            prev = st.state[key]
            hist = x[:, :, -CACHE_T:].clone()
            if hist.shape[2] < 2:
                if isinstance(prev, str):    # stream starts here: zero history (:128-129)
                    hist = torch.cat([torch.zeros_like(hist), hist], dim=2)
                else:
                    hist = torch.cat([prev[:, :, -1:], hist], dim=2)
            y = _causal_conv3d(x, sd[key + ".weight"], sd[key + ".bias"], None if isinstance(prev, str) else prev)
            st.state[key] = hist
            # channels [0:c) -> even output frames, [c:2c) -> odd (:137-139)
            y = y.reshape(b, 2, c, t, h, w)
            x = torch.stack((y[:, 0], y[:, 1]), 3).reshape(b, c, t * 2, h, w)
    if mode.startswith("upsample"):
        wgt, bias = sd[name + ".resample.1.weight"], sd[name + ".resample.1.bias"]
        x = _per_frame(x, lambda f: F.conv2d(
            F.interpolate(f.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(f), wgt, bias, padding=1))
    else:
        wgt, bias = sd[name + ".resample.1.weight"], sd[name + ".resample.1.bias"]
        x = _per_frame(x, lambda f: F.conv2d(F.pad(f, (0, 1, 0, 1)), wgt, bias, stride=2))
    if mode == "downsample3d":
        key = name + ".time_conv"
        if key not in st.state:
            st.state[key] = x.clone()        # first chunk passes through (:147-150)
        else:
            prev = st.state[key]
            last = x[:, :, -1:].clone()
            x = F.conv3d(torch.cat([prev[:, :, -1:], x], 2), sd[key + ".weight"], sd[key + ".bias"],
                         stride=(2, 1, 1))
            st.state[key] = last
    return x


def _encoder_chunk(st: _Streams, sd, cfg: VAEConfig, x: Tensor) -> Tensor:
    """Encoder3d.forward with cache (wan2pt1.py:315-357)."""
    top = cfg.dim * cfg.dim_mult[-1]
    x = _stream_conv(st, "encoder.conv1", sd, "encoder.conv1", x)
    for l in encoder_layers(cfg):
        x = _res_block(st, sd, l[1], l[2], l[3], x) if l[0] == "res" else _resample(st, sd, l[1], l[2], l[3], x)
    x = _res_block(st, sd, "encoder.middle.0", top, top, x)
    x = _attn_block(sd, "encoder.middle.1", x)
    x = _res_block(st, sd, "encoder.middle.2", top, top, x)
    x = F.silu(_rms_norm(x, sd["encoder.head.0.gamma"]))
    return _stream_conv(st, "encoder.head.2", sd, "encoder.head.2", x)


def _decoder_chunk(st: _Streams, sd, cfg: VAEConfig, x: Tensor) -> Tensor:
    """Decoder3d.forward with cache (wan2pt1.py:412-456)."""
    top = cfg.dim * cfg.dim_mult[-1]
    x = _stream_conv(st, "decoder.conv1", sd, "decoder.conv1", x)
    x = _res_block(st, sd, "decoder.middle.0", top, top, x)
    x = _attn_block(sd, "decoder.middle.1", x)
    x = _res_block(st, sd, "decoder.middle.2", top, top, x)
    for l in decoder_layers(cfg):
        x = _res_block(st, sd, l[1], l[2], l[3], x) if l[0] == "res" else _resample(st, sd, l[1], l[2], l[3], x)
    x = F.silu(_rms_norm(x, sd["decoder.head.0.gamma"]))
    return _stream_conv(st, "decoder.head.2", sd, "decoder.head.2", x)


# ----------------------------------------------------------------------------------------------
# public: what AutoencoderKLWan.encode(...).latent_dist.mode() / .decode(...)[0] compute
# ----------------------------------------------------------------------------------------------

@torch.no_grad()
def vae_encode(sd: Dict[str, Tensor], cfg: VAEConfig, x: Tensor) -> Tensor:
    """x [B,3,T,H,W] in [-1,1], T = 1 + 4k  ->  posterior mean [B,z,1+k,H/8,W/8]
    (WanVAE_.encode, wan2pt1.py:502-533, with scale = (0, 1))."""
    st = _Streams()
    t = x.shape[2]
    outs = [_encoder_chunk(st, sd, cfg, x[:, :, :1])]
    n = 1 + (t - 1) // cfg.temporal_window
    for i in range(1, n):
        outs.append(_encoder_chunk(st, sd, cfg, x[:, :, 1 + cfg.temporal_window * (i - 1): 1 + cfg.temporal_window * i]))
    if (t - 1) % cfg.temporal_window:
        outs.append(_encoder_chunk(st, sd, cfg, x[:, :, 1 + cfg.temporal_window * (n - 1):]))
    out = torch.cat(outs, 2)
    mu, _ = F.conv3d(out, sd["conv1.weight"], sd["conv1.bias"]).chunk(2, dim=1)
    return mu


@torch.no_grad()
def vae_decode(sd: Dict[str, Tensor], cfg: VAEConfig, z: Tensor, clamp: bool = True) -> Tensor:
    """z [B,z,Tl,h,w] -> video [B,3,1+4(Tl-1),8h,8w] (WanVAE_.decode, wan2pt1.py:543-560, scale = (0, 1));
    one latent frame per iteration; `clamp` = diffusers' final clamp to [-1,1]."""
    st = _Streams()
    x = F.conv3d(z, sd["conv2.weight"], sd["conv2.bias"])
    outs = [_decoder_chunk(st, sd, cfg, x[:, :, i: i + 1]) for i in range(z.shape[2])]
    out = torch.cat(outs, 2)
    return out.clamp(-1.0, 1.0) if clamp else out


def conv_flops(cfg: VAEConfig, frames_px: int, height: int, width: int, decode: bool) -> float:
    """Algorithmic conv FLOPs (2*MAC) of one encode/decode at the given PIXEL geometry, counted by walking
    the architecture with the chunk schedule above (24.58 / 41.04 TFLOP at 5x720x1280)."""
    total = 0.0
    top = cfg.dim * cfg.dim_mult[-1]

    def c3(cin, cout, k, t, h, w):
        nonlocal total
        total += 2.0 * cin * cout * k * t * h * w

    def res(cin, cout, t, h, w):
        c3(cin, cout, 27, t, h, w)
        c3(cout, cout, 27, t, h, w)
        if cin != cout:
            c3(cin, cout, 1, t, h, w)

    if decode:
        tl = 1 + (frames_px - 1) // 4
        h, w = height // 8, width // 8
        for i in range(tl):
            t, hh, ww = 1, h, w
            c3(cfg.z_dim, cfg.z_dim, 1, t, hh, ww)
            c3(cfg.z_dim, top, 27, t, hh, ww)
            res(top, top, t, hh, ww)
            total += 2.0 * top * 4 * top * t * hh * ww  # qkv + proj 1x1
            res(top, top, t, hh, ww)
            for l in decoder_layers(cfg):
                if l[0] == "res":
                    res(l[2], l[3], t, hh, ww)
                else:
                    c, mode = l[2], l[3]
                    if mode == "upsample3d" and i > 0:
                        c3(c, 2 * c, 3, t, hh, ww)
                        t *= 2
                    hh, ww = hh * 2, ww * 2
                    c3(c, c // 2, 9, t, hh, ww)
            c3(cfg.dim, 3, 27, t, hh, ww)
    else:
        chunks = [1] + [4] * ((frames_px - 1) // 4) + ([(frames_px - 1) % 4] if (frames_px - 1) % 4 else [])
        for i, t0 in enumerate(chunks):
            t, hh, ww = t0, height, width
            c3(3, cfg.dim, 27, t, hh, ww)
            for l in encoder_layers(cfg):
                if l[0] == "res":
                    res(l[2], l[3], t, hh, ww)
                else:
                    c, mode = l[2], l[3]
                    hh, ww = hh // 2, ww // 2
                    c3(c, c, 9, t, hh, ww)
                    if mode == "downsample3d" and i > 0:
                        t = (t + 1 - 3) // 2 + 1
                        c3(c, c, 3, t, hh, ww)
            res(top, top, t, hh, ww)
            total += 2.0 * top * 4 * top * t * hh * ww
            res(top, top, t, hh, ww)
            c3(top, 2 * cfg.z_dim, 27, t, hh, ww)
        tl = 1 + (frames_px - 1) // 4
        c3(2 * cfg.z_dim, 2 * cfg.z_dim, 1, tl, height // 8, width // 8)
    return total
