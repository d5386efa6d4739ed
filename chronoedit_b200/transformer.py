"""B200-native drop-in for `ChronoEditTransformer3DModel`.

Host-side mirror of /root/reference/chronoedit_diffusers/transformer_chronoedit.py:298-476: same constructor
arguments, same parameter names (so a diffusers checkpoint / `state_dict` loads unchanged), same
`forward(hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None, return_dict=True,
attention_kwargs=None)` signature and return types, `.config`, `.dtype`, `from_pretrained`.  All arithmetic happens in
libchronoedit_b200.so (hand-written sm_100a kernels) through the C ABI of include/chronoedit_b200.h; PyTorch only
owns the memory and the stream.  There is no PyTorch fallback path.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib
from ._lib import CEError, DiTConfigC, check, current_stream, ptr

# reference `_keep_in_fp32_modules` (transformer_chronoedit.py:338)
KEEP_FP32 = ("time_embedder", "scale_shift_table", "norm1", "norm2", "norm3")


@dataclass
class Transformer2DModelOutput:
    """Same shape as diffusers.models.modeling_outputs.Transformer2DModelOutput."""

    sample: torch.Tensor


class _Config(dict):
    """Attribute + item access, like diffusers' FrozenDict config."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _P(nn.Module):
    """Parameter holder (weight[, bias]) — shapes/names follow the reference module tree; no forward."""

    def __init__(self, weight_shape, bias_shape=None, dtype=torch.bfloat16, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(weight_shape, dtype=dtype, device=device), requires_grad=False)
        if bias_shape is not None:
            self.bias = nn.Parameter(torch.empty(bias_shape, dtype=dtype, device=device), requires_grad=False)


class _GELU(nn.Module):
    def __init__(self, din, dout, dtype, device):
        super().__init__()
        self.proj = _P((dout, din), (dout,), dtype, device)


class _FeedForward(nn.Module):
    def __init__(self, din, inner, dout, dtype, device):
        super().__init__()
        self.net = nn.ModuleList([_GELU(din, inner, dtype, device), nn.Identity(), _P((dout, inner), (dout,), dtype, device)])


class _Attention(nn.Module):
    def __init__(self, dim, added_kv_proj_dim, dtype, device):
        super().__init__()
        self.norm_q = _P((dim,), None, dtype, device)
        self.norm_k = _P((dim,), None, dtype, device)
        self.to_q = _P((dim, dim), (dim,), dtype, device)
        self.to_k = _P((dim, dim), (dim,), dtype, device)
        self.to_v = _P((dim, dim), (dim,), dtype, device)
        if added_kv_proj_dim is not None:
            self.add_k_proj = _P((dim, added_kv_proj_dim), (dim,), dtype, device)
            self.add_v_proj = _P((dim, added_kv_proj_dim), (dim,), dtype, device)
            self.norm_added_k = _P((dim,), None, dtype, device)
        self.to_out = nn.ModuleList([_P((dim, dim), (dim,), dtype, device), nn.Identity()])


class _Block(nn.Module):
    def __init__(self, dim, ffn_dim, added_kv_proj_dim, dtype, device):
        super().__init__()
        self.attn1 = _Attention(dim, None, dtype, device)
        self.attn2 = _Attention(dim, added_kv_proj_dim, dtype, device)
        self.norm2 = _P((dim,), (dim,), torch.float32, device)
        self.ffn = _FeedForward(dim, ffn_dim, dim, dtype, device)
        self.scale_shift_table = nn.Parameter(torch.empty(1, 6, dim, dtype=torch.float32, device=device), requires_grad=False)


class _TimestepEmbedding(nn.Module):
    def __init__(self, freq_dim, dim, device):
        super().__init__()
        self.linear_1 = _P((dim, freq_dim), (dim,), torch.float32, device)
        self.linear_2 = _P((dim, dim), (dim,), torch.float32, device)


class _TextProjection(nn.Module):
    def __init__(self, din, dim, dtype, device):
        super().__init__()
        self.linear_1 = _P((dim, din), (dim,), dtype, device)
        self.linear_2 = _P((dim, dim), (dim,), dtype, device)


class _ImageEmbedding(nn.Module):
    def __init__(self, din, dout, dtype, device):
        super().__init__()
        self.norm1 = _P((din,), (din,), torch.float32, device)
        self.ff = _FeedForward(din, din, dout, dtype, device)
        self.norm2 = _P((dout,), (dout,), torch.float32, device)


class _ConditionEmbedder(nn.Module):
    def __init__(self, dim, freq_dim, text_dim, image_dim, dtype, device):
        super().__init__()
        self.time_embedder = _TimestepEmbedding(freq_dim, dim, device)
        self.time_proj = _P((6 * dim, dim), (6 * dim,), dtype, device)
        self.text_embedder = _TextProjection(text_dim, dim, dtype, device)
        if image_dim is not None:
            self.image_embedder = _ImageEmbedding(image_dim, dim, dtype, device)


class ChronoEditTransformer3DModel(nn.Module):
    """See module docstring.  Constructor arguments = transformer_chronoedit.py:342-360."""

    config_name = "config.json"

    def __init__(
        self,
        patch_size: Tuple[int, int, int] = (1, 2, 2),
        num_attention_heads: int = 40,
        attention_head_dim: int = 128,
        in_channels: int = 16,
        out_channels: int = 16,
        text_dim: int = 4096,
        freq_dim: int = 256,
        ffn_dim: int = 13824,
        num_layers: int = 40,
        cross_attn_norm: bool = True,
        qk_norm: Optional[str] = "rms_norm_across_heads",
        eps: float = 1e-6,
        image_dim: Optional[int] = None,
        added_kv_proj_dim: Optional[int] = None,
        rope_max_seq_len: int = 1024,
        rope_temporal_skip_len: int = 8,
        *,
        torch_dtype: torch.dtype = torch.bfloat16,
        device: Optional[Union[str, torch.device]] = None,
        cache_context: bool = False,
        use_cuda_graph: bool = False,
    ) -> None:
        super().__init__()
        if torch_dtype != torch.bfloat16:
            raise CEError("chronoedit_b200 computes in bf16 (fp32 for the reference's _keep_in_fp32_modules); "
                          f"torch_dtype={torch_dtype} is not built")
        if tuple(patch_size) != (1, 2, 2) or attention_head_dim != 128:
            raise CEError("only patch_size (1,2,2) and attention_head_dim 128 are built")
        if not cross_attn_norm or qk_norm != "rms_norm_across_heads":
            raise CEError("only cross_attn_norm=True, qk_norm='rms_norm_across_heads' (the ChronoEdit configuration) is built")
        if (image_dim is None) != (added_kv_proj_dim is None):
            raise CEError("image_dim and added_kv_proj_dim must be given together (I2V configuration)")
        out_channels = out_channels or in_channels
        self.config = _Config(
            patch_size=tuple(patch_size), num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
            in_channels=in_channels, out_channels=out_channels, text_dim=text_dim, freq_dim=freq_dim, ffn_dim=ffn_dim,
            num_layers=num_layers, cross_attn_norm=cross_attn_norm, qk_norm=qk_norm, eps=eps, image_dim=image_dim,
            added_kv_proj_dim=added_kv_proj_dim, rope_max_seq_len=rope_max_seq_len,
            rope_temporal_skip_len=rope_temporal_skip_len)
        dim = num_attention_heads * attention_head_dim
        dt = torch_dtype
        self.patch_embedding = _P((dim, in_channels) + tuple(patch_size), (dim,), dt, device)
        self.condition_embedder = _ConditionEmbedder(dim, freq_dim, text_dim, image_dim, dt, device)
        self.blocks = nn.ModuleList([_Block(dim, ffn_dim, added_kv_proj_dim, dt, device) for _ in range(num_layers)])
        self.proj_out = _P((out_channels * math.prod(patch_size), dim), (out_channels * math.prod(patch_size),), dt, device)
        self.scale_shift_table = nn.Parameter(torch.empty(1, 2, dim, dtype=torch.float32, device=device), requires_grad=False)
        self._handle = None
        self._packed = False
        self._pack_keepalive: Dict[str, torch.Tensor] = {}
        self._workspaces: Dict[Tuple, torch.Tensor] = {}
        self.last_block0: Optional[torch.Tensor] = None
        # opt-in hoisting of the step-invariant context (text / image embedders, cross-attention K/V of all blocks): see
        # `ce_dit_forward_ex` in include/chronoedit_b200.h.  Up to `_ctx_slots` (prompt, negative prompt) entries, LRU.
        self.cache_context = bool(cache_context)
        # opt-in CUDA-graph replay of the forward: the ~590 launches of a forward are captured once per
        # (geometry, context slot) and replayed; inputs are staged into fixed device buffers, the result is returned from one.
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graphs: Dict[Tuple, Dict[str, Any]] = {}
        self._ctx_slots = 2
        self._ctx_cache: list = []   # entries: dict(txt=, img=, txt_v=, img_v=, shape=, buf=)
        self._lora_adapters: Dict[str, Dict[str, torch.Tensor]] = {}
        self.last_captures: Dict[int, torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------ plumbing
    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    @property
    def device(self) -> torch.device:
        return self.patch_embedding.weight.device

    def _apply(self, fn, *a, **k):
        # .to()/.cuda()/.cpu() re-allocate parameters: the fused device buffers must be rebuilt
        self._packed = False
        self._graphs = {}
        self._ctx_cache = []
        self._workspaces = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        # new weights under the same prompt tensors: the cached cross-attention K/V and any captured graph are stale; with
        # assign=True the parameters are new tensors, so the fused device buffers must be rebuilt as well
        self._ctx_cache = []
        self._graphs = {}
        if assign:
            self._packed = False
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ce_dit_destroy(self._handle)
        except Exception:
            pass

    def _cfg_c(self) -> DiTConfigC:
        c = self.config
        return DiTConfigC(
            c.num_attention_heads, c.attention_head_dim, c.in_channels, c.out_channels, c.text_dim, c.freq_dim, c.ffn_dim,
            c.num_layers, c.image_dim or 0, c.added_kv_proj_dim or 0, c.rope_max_seq_len, c.rope_temporal_skip_len, c.eps,
            *c.patch_size)

    def _ensure_handle(self):
        if self._handle is None:
            h = _lib.c_void_p()
            cfg = self._cfg_c()
            check(_lib.lib().ce_dit_create(_lib.ctypes.byref(cfg), _lib.ctypes.byref(h)))
            self._handle = h
        return self._handle

    # ------------------------------------------------------------------------------------------ fp32 validation mode
    @torch.no_grad()
    def enable_fp32_validation(self, state_dict_fp32: Dict[str, torch.Tensor]) -> None:
        """Register an fp32 copy of every parameter (reference names, e.g. the reference model's own fp32 state dict) for
        `forward_fp32`.  The bf16 parameters of this module are untouched."""
        dev = self.device
        if dev.type != "cuda":
            raise CEError("fp32 validation mode needs the module on a CUDA (sm_100) device; there is no CPU path")
        L = _lib.lib()
        self._ensure_handle()
        own = {n for n, _ in self.named_parameters()}
        missing = sorted(own - set(state_dict_fp32))
        if missing:
            raise CEError(f"enable_fp32_validation: state dict lacks {len(missing)} parameters, first: {missing[0]}")
        keep = {}
        for n in own:
            t = state_dict_fp32[n].detach().to(device=dev, dtype=torch.float32).contiguous()
            keep[n] = t
            check(L.ce_dit_set_weight_fp32(self._handle, n.encode(), ptr(t), t.numel()))
        self._fp32_keepalive = keep

    @torch.no_grad()
    def forward_fp32(self, hidden_states: torch.Tensor, timestep: torch.Tensor, encoder_hidden_states: torch.Tensor,
                     encoder_hidden_states_image: Optional[torch.Tensor] = None, return_block0: bool = False):
        """VALIDATION mode (include/chronoedit_b200.h, ce_dit_forward_fp32): fp32 in, fp32 weights, fp32 out, every matrix product
        on the tcgen05 GEMM with split-bf16 operands.  Meets rtol 1e-3 / atol 1e-4 against the reference's fp32 run end to end
        (tests/test_gpu_fp32_mode.py); far slower than forward() and not part of any benchmark."""
        if not getattr(self, "_fp32_keepalive", None):
            raise CEError("call enable_fp32_validation(fp32_state_dict) first")
        dev = self.device
        B, C, T, H, W = hidden_states.shape
        x = hidden_states.to(device=dev, dtype=torch.float32).contiguous()
        t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
        t = (t.expand(B) if t.numel() == 1 and B > 1 else t).contiguous()
        txt = encoder_hidden_states.to(device=dev, dtype=torch.float32).contiguous()
        img = None if encoder_hidden_states_image is None else encoder_hidden_states_image.to(device=dev, dtype=torch.float32).contiguous()
        L = _lib.lib()
        Lt = txt.shape[1]
        n = L.ce_dit_fp32_workspace_bytes(self._handle, B, T, H, W, Lt)
        if n < 0:
            raise CEError("fp32 validation mode: unsupported geometry")
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        out = torch.empty(B, self.config.out_channels, T, H, W, dtype=torch.float32, device=dev)
        b0 = None
        if return_block0:
            b0 = torch.empty(B * T * (H // 2) * (W // 2), self.config.num_attention_heads * self.config.attention_head_dim,
                             dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(L.ce_dit_forward_fp32(self._handle, ptr(x), ptr(t), ptr(txt), ptr(img), ptr(out), B, T, H, W, Lt, ptr(ws), ws.numel(),
                                        ptr(b0), current_stream()))
        return (out, b0) if return_block0 else out

    @torch.no_grad()
    def pack_weights(self) -> None:
        """Build the fused buffers the kernels read (QKV / KV rows stacked, scale_shift_tables stacked) and register
        every parameter with the C handle.  The reference parameters become VIEWS of the fused buffers, so no memory
        is duplicated and later in-place updates (e.g. a fused LoRA) are seen by the kernels.  Called lazily by
        forward(); call it again after replacing parameter tensors."""
        L = _lib.lib()
        dev = self.device
        if dev.type != "cuda":
            raise CEError("ChronoEditTransformer3DModel must live on a CUDA (sm_100) device; there is no CPU path")
        for n, p in self.named_parameters():
            want = torch.float32 if any(k in n for k in KEEP_FP32) else torch.bfloat16
            if p.dtype != want:
                p.data = p.data.to(want)
        self._ensure_handle()
        keep: Dict[str, torch.Tensor] = {}

        def fuse(prefix, holders, attr):
            parts = [getattr(m, attr) for m in holders]
            fused = torch.cat([p.data for p in parts], dim=0).contiguous()
            off = 0
            for p in parts:
                n0 = p.shape[0]
                p.data = fused[off: off + n0]
                off += n0
            keep[prefix + "." + attr] = fused
            return fused

        def reg(name, t):
            t = t.contiguous() if not t.is_contiguous() else t
            keep[name] = t
            check(L.ce_dit_set_weight(self._handle, name.encode(), ptr(t), 1 if t.dtype == torch.float32 else 0, t.numel()))

        fused_names = set()
        for i, blk in enumerate(self.blocks):
            p = f"blocks.{i}."
            for attr in ("weight", "bias"):
                reg(p + "attn1.to_qkv." + attr, fuse(p + "attn1.to_qkv", [blk.attn1.to_q, blk.attn1.to_k, blk.attn1.to_v], attr))
                reg(p + "attn2.to_kv." + attr, fuse(p + "attn2.to_kv", [blk.attn2.to_k, blk.attn2.to_v], attr))
                if self.config.image_dim is not None:
                    reg(p + "attn2.add_kv_proj." + attr, fuse(p + "attn2.add_kv_proj", [blk.attn2.add_k_proj, blk.attn2.add_v_proj], attr))
            fused_names.update(p + s for s in ("attn1.to_q.", "attn1.to_k.", "attn1.to_v.", "attn2.to_k.", "attn2.to_v.",
                                               "attn2.add_k_proj.", "attn2.add_v_proj."))
        table = torch.cat([blk.scale_shift_table.data for blk in self.blocks], dim=0).contiguous()  # [layers, 6, D]
        for i, blk in enumerate(self.blocks):
            blk.scale_shift_table.data = table[i: i + 1]
        reg("blocks.scale_shift_table", table)
        for n, p in self.named_parameters():
            if any(n.startswith(f) for f in fused_names) or (n.startswith("blocks.") and n.endswith("scale_shift_table")):
                continue
            reg(n, p.data)
        check(L.ce_dit_weights_complete(self._handle))
        self._pack_keepalive = keep
        self._packed = True
        self._ctx_cache = []   # cached K/V were projected with the previous weights
        self._graphs = {}

    def _workspace(self, B, T, H, W, Lt) -> torch.Tensor:
        key = (B, T, H, W, Lt, str(self.device))
        ws = self._workspaces.get(key)
        if ws is None:
            n = _lib.lib().ce_dit_workspace_bytes(self._handle, B, T, H, W, Lt)
            if n < 0:
                check(-1)
            self._workspaces.clear()  # one live geometry at a time keeps the footprint bounded
            ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._workspaces[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(
        self,
        hidden_states: torch.Tensor,
        timestep: torch.LongTensor,
        encoder_hidden_states: torch.Tensor,
        encoder_hidden_states_image: Optional[torch.Tensor] = None,
        return_dict: bool = True,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        return_block0: bool = False,
        capture_layers: Optional[Tuple[int, ...]] = None,
    ) -> Union[Transformer2DModelOutput, Tuple[torch.Tensor]]:
        # attention_kwargs["scale"] only matters for un-fused PEFT LoRA layers (transformer_chronoedit.py:406-419): without the
        # PEFT backend the reference warns and ignores it (:416-419); this mirror holds fused weights only, so it does the same.
        if attention_kwargs is not None and float(attention_kwargs.get("scale", 1.0)) != 1.0:
            import warnings

            warnings.warn("Passing `scale` via `attention_kwargs` when not using the PEFT backend is ineffective "
                          "(LoRA must be fused with fuse_lora(lora_scale=...) before inference).")
        dev = self.device
        if hidden_states.dim() != 5:
            raise CEError("hidden_states must be [B, C, T, H, W]")
        B, C, T, H, W = hidden_states.shape
        if C != self.config.in_channels:
            raise CEError(f"hidden_states has {C} channels, model expects {self.config.in_channels}")
        if (encoder_hidden_states_image is None) != (self.config.image_dim is None):
            raise CEError("encoder_hidden_states_image must be passed iff the model was built with image_dim")
        x = hidden_states.to(device=dev, dtype=torch.bfloat16).contiguous()
        t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        t = t.contiguous()
        txt = encoder_hidden_states.to(device=dev, dtype=torch.bfloat16).contiguous()
        img = None
        if encoder_hidden_states_image is not None:
            img = encoder_hidden_states_image.to(device=dev, dtype=torch.bfloat16).contiguous()
            if img.shape[1] != 257:
                raise CEError("encoder_hidden_states_image must have 257 tokens (transformer_chronoedit.py:53)")
        Lt = txt.shape[1]
        out = torch.empty(B, self.config.out_channels, T, H, W, dtype=torch.bfloat16, device=dev)
        L_tok = T * (H // 2) * (W // 2)
        Dm = self.config.num_attention_heads * self.config.attention_head_dim
        b0 = torch.empty(B * L_tok, Dm, dtype=torch.bfloat16, device=dev) if return_block0 else None
        caps: Dict[int, torch.Tensor] = {}
        if capture_layers:
            caps = {int(l): torch.empty(B * L_tok, Dm, dtype=torch.bfloat16, device=dev) for l in capture_layers}
        # (no graph replay in sequence-parallel mode: the peer barrier's epoch is a launch argument and must advance every forward)
        if self.use_cuda_graph and not caps and b0 is None and not getattr(self, "_sp_state", None):
            out = self._graphed_forward(x, t, txt, img, encoder_hidden_states, encoder_hidden_states_image)
        else:
            self._native_forward(x, t, txt, img, out, b0, caps, encoder_hidden_states, encoder_hidden_states_image)
        self.last_block0 = b0
        self.last_captures = caps
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)

    def _native_forward(self, x, t, txt, img, out, b0, caps, txt_in, img_in) -> None:
        """The one place the DiT forward reaches the C ABI (ce_dit_forward_ex).  x [B,C,T,H,W] bf16, t [B] fp32, txt / img bf16,
        all contiguous on the module's device; `out` (and the optional block captures) are written in place."""
        if not self._packed:
            self.pack_weights()
        B, _, T, H, W = x.shape
        Lt = txt.shape[1]
        L = _lib.lib()
        ws = self._workspace(B, T, H, W, Lt)
        if caps:
            lay = (_lib.c_int32 * len(caps))(*caps.keys())
            dst = (_lib.c_void_p * len(caps))(*[v.data_ptr() for v in caps.values()])
            check(L.ce_dit_set_capture(self._handle, lay, dst, len(caps)))
        ctx_buf, ctx_reuse = (None, 0)
        if self.cache_context:
            ctx_buf, ctx_reuse = self._context_slot(txt_in, img_in, txt, img)
        try:
            with torch.cuda.device(self.device):
                check(L.ce_dit_forward_ex(self._handle, ptr(x), ptr(t), ptr(txt), ptr(img), ptr(out), B, T, H, W, Lt, ptr(ws), ws.numel(),
                                          ptr(b0), ptr(ctx_buf), ctx_buf.numel() if ctx_buf is not None else 0, ctx_reuse,
                                          current_stream()))
        finally:
            if caps:
                check(L.ce_dit_set_capture(self._handle, None, None, 0))

    def _graphed_forward(self, x, t, txt, img, txt_in, img_in) -> torch.Tensor:
        """Replay (or, the second time a configuration is seen, capture) the forward as one CUDA graph.  A configuration = geometry +
        the context-cache slot in use (its buffer address is baked into the captured launches).  The first call of a configuration
        runs eagerly: it packs weights, builds RoPE tables, opts kernels into their shared-memory sizes and fills the context cache
        -- none of which may happen during capture.  Inputs are copied into fixed staging tensors; the returned sample is a fresh
        tensor (a copy out of the graph's fixed output buffer)."""
        if not self._packed:
            self.pack_weights()
        ctx_buf, ctx_reuse = (None, 0)
        if self.cache_context:
            ctx_buf, ctx_reuse = self._context_slot(txt_in, img_in, txt, img)
        key = (tuple(x.shape), tuple(txt.shape), None if img is None else tuple(img.shape), None if ctx_buf is None else ctx_buf.data_ptr())
        L = _lib.lib()
        B, _, T, H, W = x.shape
        Lt = txt.shape[1]

        def launch(xs, ts, txts, imgs, outs, reuse):
            ws = self._workspace(B, T, H, W, Lt)
            with torch.cuda.device(self.device):
                check(L.ce_dit_forward_ex(self._handle, ptr(xs), ptr(ts), ptr(txts), ptr(imgs), ptr(outs), B, T, H, W, Lt, ptr(ws), ws.numel(),
                                          None, ptr(ctx_buf), ctx_buf.numel() if ctx_buf is not None else 0, reuse, current_stream()))

        g = self._graphs.get(key)
        if g is None or (self.cache_context and not ctx_reuse):
            # eager: first sight of this configuration, or the context has to be (re)computed into the slot
            out = torch.empty(B, self.config.out_channels, T, H, W, dtype=torch.bfloat16, device=self.device)
            launch(x, t, txt, img, out, ctx_reuse)
            if g is None:
                self._graphs[key] = {"graph": None}
            return out
        if g["graph"] is None:   # second sight: capture (with ctx_reuse = 1 when the context cache is on)
            st = {"x": x.clone(), "t": t.clone(), "txt": txt.clone(), "img": None if img is None else img.clone(),
                  "out": torch.empty(B, self.config.out_channels, T, H, W, dtype=torch.bfloat16, device=self.device)}
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                launch(st["x"], st["t"], st["txt"], st["img"], st["out"], 1 if self.cache_context else 0)
            g.update(st)
            g["graph"] = graph
        g["x"].copy_(x)
        g["t"].copy_(t)
        if not self.cache_context:   # with the context cache the captured launches do not read the encoder states at all
            g["txt"].copy_(txt)
            if img is not None:
                g["img"].copy_(img)
        g["graph"].replay()
        return g["out"].clone()

    def _context_slot(self, txt_in, img_in, txt, img):
        """(cache buffer, reuse flag) for these encoder states.  A hit requires the SAME tensor objects the cache entry was
        filled from (the entry holds references, so their storage cannot have been recycled) at the same in-place version
        counter -- the unchanged pipeline passes the same `prompt_embeds` / `negative_prompt_embeds` / `image_embeds` objects at
        every step (pipeline_chronoedit.py:715-735)."""
        ver = lambda t: None if t is None else t._version
        for e in self._ctx_cache:
            if e["txt"] is txt_in and e["img"] is img_in and e["txt_v"] == ver(txt_in) and e["img_v"] == ver(img_in):
                self._ctx_cache.remove(e)
                self._ctx_cache.append(e)   # most recently used last
                return e["buf"], 1
        B, Lt = txt.shape[0], txt.shape[1]
        n = _lib.lib().ce_dit_context_cache_bytes(self._handle, B, Lt)
        if n < 0:
            check(-1)
        buf = None
        if len(self._ctx_cache) >= self._ctx_slots:
            old = self._ctx_cache.pop(0)
            if old["buf"].numel() >= n and old["buf"].device == self.device:
                buf = old["buf"]
        if buf is None:
            buf = torch.empty(n, dtype=torch.uint8, device=self.device)
        self._ctx_cache.append(dict(txt=txt_in, img=img_in, txt_v=ver(txt_in), img_v=ver(img_in), buf=buf))
        return buf, 0

    def clear_context_cache(self) -> None:
        self._ctx_cache = []

    @torch.no_grad()
    def forward_host(self, hidden_states: torch.Tensor, timestep: torch.Tensor, encoder_hidden_states: torch.Tensor,
                     encoder_hidden_states_image: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                     return_device_sample: bool = False):
        """End-to-end variant: all tensors are (pinned) HOST bf16/fp32 tensors; H2D, forward and D2H happen inside the
        C-ABI call (ce_dit_forward_host_ex).  Returns the host sample tensor (and, with `return_device_sample`, a view of the same
        sample still on the device -- valid until the next forward_host call -- for the scheduler step)."""
        if not self._packed:
            self.pack_weights()
        B, C, T, H, W = hidden_states.shape
        Lt = encoder_hidden_states.shape[1]
        for tns in (hidden_states, encoder_hidden_states):
            if tns.device.type != "cpu" or tns.dtype != torch.bfloat16 or not tns.is_contiguous():
                raise CEError("forward_host expects contiguous CPU bf16 tensors")
        t = timestep.to(torch.float32).contiguous()
        L = _lib.lib()
        ws = self._workspace(B, T, H, W, Lt)
        n = L.ce_dit_host_staging_bytes(self._handle, B, T, H, W, Lt)
        st = self._workspaces.get("staging")
        if st is None or st.numel() < n:
            st = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._workspaces["staging"] = st
        if out is None:
            out = torch.empty(B, self.config.out_channels, T, H, W, dtype=torch.bfloat16).pin_memory()
        ctx_buf, ctx_reuse = (None, 0)
        if self.cache_context:
            ctx_buf, ctx_reuse = self._context_slot(encoder_hidden_states, encoder_hidden_states_image, encoder_hidden_states,
                                                    encoder_hidden_states_image)
        with torch.cuda.device(self.device):
            check(L.ce_dit_forward_host_ex(self._handle, ptr(hidden_states), ptr(t), ptr(encoder_hidden_states),
                                           ptr(encoder_hidden_states_image), ptr(out), B, T, H, W, Lt, ptr(st), st.numel(), ptr(ws),
                                           ws.numel(), ptr(ctx_buf), ctx_buf.numel() if ctx_buf is not None else 0, ctx_reuse,
                                           current_stream()))
        if not return_device_sample:
            return out
        off = L.ce_dit_host_staging_sample_offset(self._handle, B, T, H, W, Lt)
        dev_sample = st[off: off + out.numel() * 2].view(torch.bfloat16).view(out.shape)
        return out, dev_sample

    def launches_per_forward(self) -> int:
        return int(_lib.lib().ce_dit_last_launch_count(self._handle)) if self._handle else 0

    # ------------------------------------------------------------------------------------------ LoRA
    # (original Wan module name -> diffusers name): the pairs of the reference's own converters -- per-block modules
    # chronoedit/_src/models/utils.py:112-212, non-block modules :214-290; the diffusers-side names are those of
    # chronoedit_diffsynth/wan_video_dit_chronoedit.py:439-496 read right to left.
    _WAN_TO_DIFFUSERS = (("self_attn.q", "attn1.to_q"), ("self_attn.k", "attn1.to_k"), ("self_attn.v", "attn1.to_v"),
                         ("self_attn.o", "attn1.to_out.0"), ("cross_attn.q", "attn2.to_q"), ("cross_attn.k_img", "attn2.add_k_proj"),
                         ("cross_attn.v_img", "attn2.add_v_proj"), ("cross_attn.k", "attn2.to_k"), ("cross_attn.v", "attn2.to_v"),
                         ("cross_attn.o", "attn2.to_out.0"), ("ffn.0", "ffn.net.0.proj"), ("ffn.2", "ffn.net.2"))
    _WAN_TO_DIFFUSERS_GLOBAL = {
        "time_embedding.0": "condition_embedder.time_embedder.linear_1", "time_embedding.2": "condition_embedder.time_embedder.linear_2",
        "text_embedding.0": "condition_embedder.text_embedder.linear_1", "text_embedding.2": "condition_embedder.text_embedder.linear_2",
        "time_projection.1": "condition_embedder.time_proj", "head.head": "proj_out",
        "img_emb.proj.1": "condition_embedder.image_embedder.ff.net.0.proj", "img_emb.proj.3": "condition_embedder.image_embedder.ff.net.2",
    }

    def _parse_lora(self, lora_state_dict: Dict[str, torch.Tensor], adapter_name: Optional[str]):
        """Pass 1 of fuse_lora: key grammar, name mapping, presence of both factors, shapes.  Nothing is modified here, so a bad
        file leaves the model untouched.  Returns [(parameter, A, B, alpha or None)]."""
        params = dict(self.named_parameters())
        pairs: Dict[str, Dict[str, torch.Tensor]] = {}
        for key, val in lora_state_dict.items():
            k = key
            for pre in ("transformer.", "diffusion_model."):
                if k.startswith(pre):
                    k = k[len(pre):]
            if k.endswith((".diff", ".diff_b")) or ".lora_B." in k and k.endswith(".bias"):
                raise CEError(f"fuse_lora: '{key}' is a weight/bias delta, not a low-rank pair; not supported")
            k = k.replace(".lora_down.", ".lora_A.").replace(".lora_up.", ".lora_B.")
            kind = None
            for tag, nm in ((".lora_A.", "A"), (".lora_B.", "B")):
                if tag in k and k.endswith(".weight"):
                    mod, rest = k.split(tag, 1)
                    seg = rest[: -len("weight")].strip(".")       # "" or the PEFT adapter segment ("default", ...)
                    if seg and adapter_name is not None and seg != adapter_name:
                        kind = "skip"   # another adapter's weights
                    elif "." in seg:
                        raise CEError(f"fuse_lora: unrecognised key '{key}'")
                    else:
                        kind = nm
                    break
            if kind is None and k.endswith(".alpha"):
                mod, kind = k[: -len(".alpha")], "alpha"
            if kind is None:
                raise CEError(f"fuse_lora: unrecognised key '{key}'")
            if kind == "skip":
                continue
            if mod in self._WAN_TO_DIFFUSERS_GLOBAL:
                mod = self._WAN_TO_DIFFUSERS_GLOBAL[mod]
            else:
                for wan, dif in self._WAN_TO_DIFFUSERS:
                    if mod.endswith("." + wan):
                        mod = mod[: -len(wan)] + dif
                        break
            if kind in pairs.setdefault(mod, {}):
                raise CEError(f"fuse_lora: '{mod}' has more than one {kind} entry (several adapters in one file? pass adapter_name)")
            pairs[mod][kind] = val
        plan = []
        for mod, d in pairs.items():
            if "A" not in d or "B" not in d:
                raise CEError(f"fuse_lora: '{mod}' needs both lora_A and lora_B")
            w = params.get(mod + ".weight")
            if w is None:
                raise CEError(f"fuse_lora: no parameter '{mod}.weight' in this model")
            A, B = d["A"], d["B"]
            r = A.shape[0]
            if A.dim() != 2 or B.dim() != 2 or A.shape != (r, w.shape[1]) or B.shape != (w.shape[0], r):
                raise CEError(f"fuse_lora: shapes of '{mod}' do not match: A {tuple(A.shape)} B {tuple(B.shape)} W {tuple(w.shape)}")
            plan.append((w, A, B, float(d["alpha"]) if "alpha" in d else None))
        return plan

    def load_lora_adapter(self, state_dict: Dict[str, torch.Tensor], prefix: Optional[str] = "transformer", network_alphas=None,
                          adapter_name: Optional[str] = None, **unused) -> None:
        """diffusers `PeftAdapterMixin.load_lora_adapter` as `WanLoraLoaderMixin.load_lora_weights` calls it
        (run_inference_diffusers.py:371): keep the adapter's factors (validated now) until `fuse_lora` merges them."""
        if prefix:
            sub = {k[len(prefix) + 1:]: v for k, v in state_dict.items() if k.startswith(prefix + ".")}
            state_dict = sub or {k: v for k, v in state_dict.items() if not k.startswith(("text_encoder.", "vae."))}
        name = adapter_name or f"default_{len(self._lora_adapters)}"
        self._parse_lora(state_dict, None)   # fail now, not at fuse time
        self._lora_adapters[name] = dict(state_dict)

    def unload_lora(self) -> None:
        self._lora_adapters.clear()

    @torch.no_grad()
    def fuse_lora(self, *args, **kwargs) -> int:
        """`pipe.load_lora_weights(path); pipe.fuse_lora(lora_scale=s)` (run_inference_diffusers.py:369-376) for this module:
        W += (B @ A) * (s * alpha / r), in place and in the weight dtype, which is PEFT's merge arithmetic.  The kernels see
        the result without repacking because the fused QKV buffers are views of the same storage.

        Two call styles: diffusers' `fuse_lora(lora_scale, safe_fusing=False, adapter_names=None)` merges the adapters handed over
        by `load_lora_adapter`; `fuse_lora(state_dict, lora_scale=s)` merges a state dict directly.  Keys: diffusers / PEFT style
        `[transformer.]blocks.N.attn1.to_q.lora_A[.adapter].weight` (+ `lora_B`, optional `.alpha`), or the original Wan style the
        in-tree loader converts (`[diffusion_model.]blocks.N.self_attn.q.lora_down|lora_A.weight`, `lora_up|lora_B`, `.alpha`, and
        the non-block modules time_embedding / text_embedding / time_projection / head.head / img_emb.proj;
        chronoedit/_src/models/utils.py:66-290, wan_t2v_model.py:385-391).  `diff` / `diff_b` entries (norm / bias deltas the
        reference converter drops or treats as lora_bias) are rejected.  Everything is validated BEFORE the first weight is
        touched.  One-time weight preparation with torch matmuls; not on the per-step path.  Returns the number of weights updated."""
        lora_state_dict = kwargs.pop("lora_state_dict", None)
        if args and isinstance(args[0], dict):        # fuse_lora(state_dict[, lora_scale])
            lora_state_dict, args = args[0], args[1:]
        lora_scale = kwargs.pop("lora_scale", args[0] if len(args) > 0 else 1.0)
        safe_fusing = bool(kwargs.pop("safe_fusing", args[1] if len(args) > 1 else False))
        adapter_names = kwargs.pop("adapter_names", args[2] if len(args) > 2 else None)
        adapter_name = kwargs.pop("adapter_name", None)
        if kwargs:
            raise TypeError(f"fuse_lora: unexpected arguments {sorted(kwargs)}")
        if lora_state_dict is not None:
            plans = [self._parse_lora(lora_state_dict, adapter_name)]
        else:
            names = list(self._lora_adapters) if adapter_names is None else list(adapter_names)
            missing = [n for n in names if n not in self._lora_adapters]
            if missing or not names:
                raise CEError(f"fuse_lora: no loaded adapter named {missing or '(none loaded)'}")
            plans = [self._parse_lora(self._lora_adapters[n], None) for n in names]
        n = 0
        deltas = []
        for plan in plans:
            for w, A, B, alpha in plan:
                A, B = A.to(w.device, w.dtype), B.to(w.device, w.dtype)
                r = A.shape[0]
                delta = (B @ A) * (float(lora_scale) * (alpha if alpha is not None else float(r)) / r)   # peft get_delta_weight
                if safe_fusing and not torch.isfinite(delta).all():
                    raise CEError("fuse_lora(safe_fusing=True): non-finite values in the LoRA delta; nothing was fused")
                deltas.append((w, delta))
        for w, delta in deltas:
            w.data += delta
            n += 1
        if lora_state_dict is None:
            self._lora_adapters.clear()   # merged: a second fuse_lora must not add them again
        self._ctx_cache = []              # cross-attention K/V cached from the old weights are stale
        return n

    # ------------------------------------------------------------------------------------------ loading
    @classmethod
    def from_config(cls, config: Dict[str, Any], **kw) -> "ChronoEditTransformer3DModel":
        fields = ("patch_size", "num_attention_heads", "attention_head_dim", "in_channels", "out_channels", "text_dim", "freq_dim",
                  "ffn_dim", "num_layers", "cross_attn_norm", "qk_norm", "eps", "image_dim", "added_kv_proj_dim",
                  "rope_max_seq_len", "rope_temporal_skip_len")
        return cls(**{k: config[k] for k in fields if k in config}, **kw)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None,
                        torch_dtype: torch.dtype = torch.bfloat16, device: Optional[Union[str, torch.device]] = None,
                        **unused) -> "ChronoEditTransformer3DModel":
        """Load a diffusers-format checkpoint directory (config.json + diffusion_pytorch_model*.safetensors), the layout
        `run_inference_diffusers.py:349-353` reads.  Local paths only (no hub download)."""
        from safetensors import safe_open

        root = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        with open(os.path.join(root, cls.config_name)) as f:
            cfg = json.load(f)
        model = cls.from_config(cfg, torch_dtype=torch_dtype, device=device or "cpu")
        index = os.path.join(root, "diffusion_pytorch_model.safetensors.index.json")
        if os.path.exists(index):
            with open(index) as f:
                files = sorted(set(json.load(f)["weight_map"].values()))
        else:
            files = ["diffusion_pytorch_model.safetensors"]
        params = dict(model.named_parameters())
        seen = set()
        for fn in files:
            with safe_open(os.path.join(root, fn), framework="pt", device="cpu") as sf:
                for k in sf.keys():
                    if "norm_added_q" in k:  # `_keys_to_ignore_on_load_unexpected` (transformer_chronoedit.py:339)
                        continue
                    if k not in params:
                        raise CEError(f"unexpected key in checkpoint: {k}")
                    params[k].data.copy_(sf.get_tensor(k).reshape(params[k].shape))
                    seen.add(k)
        missing = sorted(set(params) - seen)
        if missing:
            raise CEError(f"checkpoint is missing {len(missing)} parameters, first: {missing[0]}")
        return model
