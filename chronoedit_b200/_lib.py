"""ctypes binding of libchronoedit_b200.so (the C ABI declared in include/chronoedit_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing this module raises, and every compute
entry point fails on a machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libchronoedit_b200.so")


class CEError(RuntimeError):
    pass


class VAEConfigC(ctypes.Structure):
    _fields_ = [("dim", c_int32), ("z_dim", c_int32), ("dim_mult", c_int32 * 4), ("num_res_blocks", c_int32),
                ("temporal_downsample", c_int32 * 3)]


class DiTConfigC(ctypes.Structure):
    _fields_ = [
        ("num_attention_heads", c_int32), ("attention_head_dim", c_int32), ("in_channels", c_int32),
        ("out_channels", c_int32), ("text_dim", c_int32), ("freq_dim", c_int32), ("ffn_dim", c_int32),
        ("num_layers", c_int32), ("image_dim", c_int32), ("added_kv_proj_dim", c_int32),
        ("rope_max_seq_len", c_int32), ("rope_temporal_skip_len", c_int32), ("eps", c_float),
        ("patch_t", c_int32), ("patch_h", c_int32), ("patch_w", c_int32),
    ]


class EncoderConfigC(ctypes.Structure):   # ce_encoder_config
    _fields_ = [("kind", c_int32), ("vocab_size", c_int32), ("d_model", c_int32), ("d_kv", c_int32), ("d_ff", c_int32),
                ("num_layers", c_int32), ("num_heads", c_int32), ("eps", c_float), ("image_size", c_int32), ("patch_size", c_int32),
                ("hidden_act", c_int32)]


class UniPCStepArgsC(ctypes.Structure):   # ce_unipc_step_args (field order = include/chronoedit_b200.h)
    _fields_ = [
        ("sample_dtype", c_int32), ("model_dtype", c_int32), ("n", c_int64), ("cond", c_void_p), ("uncond", c_void_p),
        ("guidance", c_float), ("sigma", c_float), ("sample", c_void_p), ("last_sample", c_void_p), ("m_prev", c_void_p),
        ("m_prev2", c_void_p), ("use_corrector", c_int32), ("c_order", c_int32), ("c_x", c_float), ("c_m0", c_float),
        ("c_bh", c_float), ("c_inv_rk", c_float), ("c_rho0", c_float), ("c_rho1", c_float), ("p_order", c_int32),
        ("p_x", c_float), ("p_m0", c_float), ("p_bh", c_float), ("p_inv_rk", c_float), ("p_zero", c_float),
        ("x0_out", c_void_p), ("corrected_out", c_void_p), ("prev_sample_out", c_void_p), ("model_input_out", c_void_p),
        ("inner", c_int64), ("c_lat", c_int32), ("c_total", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/chronoedit_b200.h declares
SIGNATURES = {
    "ce_abi_version": (c_int, []),
    "ce_last_error": (c_char_p, []),
    "ce_device_check": (c_int, []),
    "ce_unipc_step": (c_int, [POINTER(UniPCStepArgsC), c_void_p]),
    "ce_dit_create": (c_int, [POINTER(DiTConfigC), POINTER(c_void_p)]),
    "ce_dit_destroy": (None, [c_void_p]),
    "ce_dit_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int64]),
    "ce_dit_weights_complete": (c_int, [c_void_p]),
    "ce_dit_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ce_dit_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "ce_dit_context_cache_bytes": (c_int64, [c_void_p, c_int, c_int]),
    "ce_dit_forward_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "ce_dit_sp_region_bytes": (c_int64, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ce_dit_sp_configure": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64]),
    "ce_ipc_alloc": (c_int, [c_int64, POINTER(c_void_p)]),
    "ce_ipc_free": (c_int, [c_void_p]),
    "ce_ipc_get_handle": (c_int, [c_void_p, c_void_p]),
    "ce_ipc_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "ce_ipc_close": (c_int, [c_void_p]),
    "ce_dit_set_capture": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "ce_dit_host_staging_bytes": (c_int64, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ce_dit_forward_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "ce_dit_forward_host_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "ce_dit_host_staging_sample_offset": (c_int64, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ce_dit_set_weight_fp32": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "ce_dit_fp32_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ce_dit_forward_fp32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_int64, c_void_p, c_void_p]),
    "ce_dit_last_launch_count": (c_int64, [c_void_p]),
    "ce_dit_profile_begin": (c_int, [c_void_p, c_int]),
    "ce_dit_profile_end": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ce_encoder_create": (c_int, [POINTER(EncoderConfigC), POINTER(c_void_p)]),
    "ce_encoder_destroy": (None, [c_void_p]),
    "ce_encoder_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "ce_encoder_workspace_bytes": (c_int64, [c_void_p, c_int, c_int]),
    "ce_encoder_last_launch_count": (c_int64, [c_void_p]),
    "ce_umt5_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "ce_clip_vision_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "ce_vae_create": (c_int, [c_void_p, POINTER(c_void_p)]),
    "ce_vae_destroy": (None, [c_void_p]),
    "ce_vae_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "ce_vae_workspace_bytes": (c_int64, [c_void_p, c_int, c_int, c_int, c_int]),
    "ce_vae_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "ce_vae_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "ce_vae_last_launch_count": (c_int64, [c_void_p]),
    "ce_conv3d_cl_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ce_linear_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                               c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "ce_attention_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                  c_int, c_int, c_float, c_int, c_void_p]),
    "ce_debug_attention_kernel": (c_int, [c_int]),
    "ce_debug_attention_timing": (c_int, [c_void_p]),
    "ce_attention_dual_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "ce_layernorm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int,
                                  c_int, c_void_p, c_void_p, c_void_p]),
    "ce_rmsnorm_rope_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int, c_void_p]),
    "ce_rope_table_host": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library with typed signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CEError(
                f"{LIB_PATH} is missing: build it with `python -m chronoedit_b200.build` (or __graft_entry__.build()). "
                "chronoedit_b200 has no CPU / PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        if l.ce_abi_version() != 1:
            raise CEError(f"ABI version mismatch: library reports {l.ce_abi_version()}, binding expects 1")
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().ce_last_error()
        raise CEError(f"chronoedit_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t) -> c_void_p:
    """Device (or host) data pointer of a torch tensor as c_void_p; None -> NULL."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def current_stream() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
