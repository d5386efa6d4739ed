"""Attention container (diffusers/models/attention_processor.py, 0.35.2) — only
the members ChronoEditAttnProcessor2_0 touches (transformer_chronoedit.py:43-108)."""
import inspect

import torch.nn as nn

from .normalization import RMSNorm


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, kv_heads=None, dim_head=64,
                 dropout=0.0, bias=False, qk_norm=None, added_kv_proj_dim=None, added_proj_bias=True,
                 out_bias=True, eps=1e-5, processor=None, out_dim=None):
        super().__init__()
        self.inner_dim = out_dim if out_dim is not None else dim_head * heads
        self.inner_kv_dim = self.inner_dim if kv_heads is None else dim_head * kv_heads
        self.query_dim = query_dim
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.out_dim = out_dim if out_dim is not None else query_dim
        self.heads = heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.scale = dim_head ** -0.5

        if qk_norm is None:
            self.norm_q = None
            self.norm_k = None
        elif qk_norm == "rms_norm_across_heads":
            self.norm_q = RMSNorm(dim_head * heads, eps=eps)
            self.norm_k = RMSNorm(dim_head * (kv_heads or heads), eps=eps)
        else:  # pragma: no cover
            raise NotImplementedError(qk_norm)

        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_kv_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_kv_dim, bias=bias)

        self.add_k_proj = None
        self.add_v_proj = None
        self.norm_added_q = None
        self.norm_added_k = None
        if added_kv_proj_dim is not None:
            self.add_k_proj = nn.Linear(added_kv_proj_dim, self.inner_kv_dim, bias=added_proj_bias)
            self.add_v_proj = nn.Linear(added_kv_proj_dim, self.inner_kv_dim, bias=added_proj_bias)
            if qk_norm == "rms_norm_across_heads":
                # Wan: norm across all heads, no q-norm on the added stream
                self.norm_added_k = RMSNorm(dim_head * (kv_heads or heads), eps=eps)

        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, self.out_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        accepted = set(inspect.signature(self.processor.__call__).parameters.keys())
        kwargs = {k: v for k, v in kwargs.items() if k in accepted}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)
