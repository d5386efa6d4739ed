#!/usr/bin/env python
"""Operator micro-benchmarks at the 14B / 720p shapes (CUDA events, warm-up, L2-busting operand rotation).
    python scripts/bench_ops.py [attn] [gemm] [rows] [conv]
Prints one JSON line per op; used to iterate on a kernel without running the whole 40-layer step."""
import json
import math
import sys
import time

import torch

sys.path.insert(0, ".")
import chronoedit_b200._lib as L  # noqa: E402

lib = L.lib()


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_attn(B=2, H=40, Lq=7200, Lk=7200, nbuf=3):
    D = H * 128
    bufs = [torch.randn(B, Lq, 3 * D, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)] if Lq == Lk else None
    out = torch.empty(B, Lq, D, device="cuda", dtype=torch.bfloat16)
    i = [0]
    if bufs is None:
        q = torch.randn(B, Lq, D, device="cuda", dtype=torch.bfloat16)
        kv = torch.randn(B, Lk, 2 * D, device="cuda", dtype=torch.bfloat16)

    def fn():
        if bufs is not None:
            x = bufs[i[0] % nbuf]
            i[0] += 1
            L.check(lib.ce_attention_bf16(L.ptr(x), 3 * D, L.ptr(x[..., D:]), 3 * D, L.ptr(x[..., 2 * D:]), 3 * D, L.ptr(out), D, B, H, Lq, Lk,
                                          1 / math.sqrt(128), 0, L.current_stream()))
        else:
            L.check(lib.ce_attention_bf16(L.ptr(q), D, L.ptr(kv), 2 * D, L.ptr(kv[..., D:]), 2 * D, L.ptr(out), D, B, H, Lq, Lk,
                                          1 / math.sqrt(128), 0, L.current_stream()))

    ms = timeit(fn)
    fl = 4.0 * B * H * Lq * Lk * 128
    print(json.dumps({"op": "attention", "B": B, "H": H, "Lq": Lq, "Lk": Lk, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)


def bench_attn_libs(B=2, H=40, L=7200, Lk=None):
    """The library bar for the self-attention: what the unmodified reference executes on a cc-10.0 GPU --
    torch SDPA with the cuDNN fused-attention backend (chronoedit/_src/modules/attention.py:129-138 picks it on Blackwell), torch's
    built-in flash backend, and the flash-attn 2.8 package -- on the same shapes, same timing loop."""
    from torch.nn.attention import SDPBackend, sdpa_kernel

    Lk = Lk or L
    q = torch.randn(B, H, L, 128, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, H, Lk, 128, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, H, Lk, 128, device="cuda", dtype=torch.bfloat16)
    fl = 4.0 * B * H * L * Lk * 128
    for name, backend in (("sdpa_cudnn", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash", SDPBackend.FLASH_ATTENTION),
                          ("sdpa_efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel(backend):
                ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters=5, warmup=2)
            print(json.dumps({"op": name, "B": B, "H": H, "Lq": L, "Lk": Lk, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"op": name, "Lq": L, "error": str(e)[:160]}), flush=True)
    try:
        from flash_attn import flash_attn_func

        qf, kf, vf = (t.transpose(1, 2).contiguous() for t in (q, k, v))
        ms = timeit(lambda: flash_attn_func(qf, kf, vf), iters=5, warmup=2)
        print(json.dumps({"op": "flash_attn_2.8_pkg", "B": B, "H": H, "Lq": L, "Lk": Lk, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"op": "flash_attn_2.8_pkg", "error": str(e)[:160]}), flush=True)


def bench_attn_dual(B=2, H=40, Lq=7200, Lk=512, Lk2=257):
    """The cross-attention launch of the DiT block: text keys + image keys in one kernel."""
    D = H * 128
    q = torch.randn(B, Lq, D, device="cuda", dtype=torch.bfloat16)
    kv = torch.randn(B, Lk, 2 * D, device="cuda", dtype=torch.bfloat16)
    kv2 = torch.randn(B, Lk2, 2 * D, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(B, Lq, D, device="cuda", dtype=torch.bfloat16)

    def fn():
        L.check(lib.ce_attention_dual_bf16(L.ptr(q), D, L.ptr(kv), 2 * D, L.ptr(kv[..., D:]), 2 * D, L.ptr(kv2), 2 * D, L.ptr(kv2[..., D:]), 2 * D,
                                           L.ptr(out), D, B, H, Lq, Lk, Lk2, 1 / math.sqrt(128), L.current_stream()))

    ms = timeit(fn)
    fl = 4.0 * B * H * Lq * (Lk + Lk2) * 128
    print(json.dumps({"op": "attention_dual", "B": B, "H": H, "Lq": Lq, "Lk": Lk, "Lk2": Lk2, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)


def bench_gemm(M, N, K, epi=0, nbuf=3):
    A = [torch.randn(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    W = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(nbuf)]
    bias = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    resid = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    gate = torch.randn(2, N, device="cuda", dtype=torch.float32)
    i = [0]

    def fn():
        a, w = A[i[0] % nbuf], W[i[0] % nbuf]
        i[0] += 1
        L.check(lib.ce_linear_bf16(L.ptr(a), K, L.ptr(w), K, L.ptr(bias), L.ptr(out), N, None, M, N, K, epi, L.ptr(resid), N, L.ptr(gate), N,
                                   (M + 1) // 2, L.current_stream()))

    ms = timeit(fn)
    print(json.dumps({"op": "gemm", "M": M, "N": N, "K": K, "epi": epi, "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}), flush=True)
    # library bar: cuBLAS through torch on the same shapes
    ms2 = timeit(lambda: torch.matmul(A[0], W[0].t()))
    print(json.dumps({"op": "cublas_gemm", "M": M, "N": N, "K": K, "ms": round(ms2, 4), "tflops": round(2.0 * M * N * K / ms2 / 1e9, 1)}), flush=True)


def bench_rows(M=14400, D=5120):
    x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    y = torch.empty_like(x)
    mod = torch.randn(2, 6, D, device="cuda")
    ms = timeit(lambda: L.check(lib.ce_layernorm_bf16(L.ptr(x), D, L.ptr(y), D, M, D, 1e-6, L.ptr(mod[:, 1]), L.ptr(mod[:, 0]), 6 * D, M // 2, None,
                                                      None, L.current_stream())))
    print(json.dumps({"op": "layernorm_mod", "M": M, "D": D, "ms": round(ms, 4), "GBps": round(4.0 * M * D / ms / 1e6, 1)}), flush=True)
    w = torch.ones(D, device="cuda", dtype=torch.bfloat16)
    cos = torch.randn(M // 2, 64, device="cuda")
    ms = timeit(lambda: L.check(lib.ce_rmsnorm_rope_bf16(L.ptr(x), D, M, D, 1e-6, L.ptr(w), L.ptr(cos), L.ptr(cos), M // 2, 128, L.current_stream())))
    print(json.dumps({"op": "rmsnorm_rope", "M": M, "D": D, "ms": round(ms, 4), "GBps": round(4.0 * M * D / ms / 1e6, 1)}), flush=True)


def bench_conv(Cin=96, Cout=96, T=4, H=720, W=1280):
    from chronoedit_b200.autoencoder import pack_parameter

    x = torch.randn(T + 2, H, W, Cin, device="cuda", dtype=torch.bfloat16)
    w = pack_parameter("l.weight", torch.randn(Cout, Cin, 3, 3, 3, device="cuda") * 0.02)
    b = torch.zeros(Cout, device="cuda", dtype=torch.bfloat16)
    y = torch.empty(T, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: L.check(lib.ce_conv3d_cl_bf16(L.ptr(x), T + 2, H, W, Cin, L.ptr(w), L.ptr(b), Cout, 3, 3, 3, 1, 1, 1, 1, 1, 0, L.ptr(y), T, H, W,
                                                      None, 0, L.current_stream())), iters=5, warmup=2)
    fl = 2.0 * 27 * Cin * Cout * T * H * W
    print(json.dumps({"op": "conv3x3x3", "Cin": Cin, "Cout": Cout, "T": T, "H": H, "W": W, "ms": round(ms, 3), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
    xt = x.permute(3, 0, 1, 2)[None].contiguous(memory_format=torch.channels_last_3d)
    wt = torch.randn(Cout, Cin, 3, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    try:
        ms2 = timeit(lambda: torch.nn.functional.conv3d(xt, wt, padding=(0, 1, 1)), iters=3, warmup=1)
        print(json.dumps({"op": "cudnn_conv3d", "ms": round(ms2, 3), "tflops": round(fl / ms2 / 1e9, 1)}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"op": "cudnn_conv3d", "error": str(e)[:200]}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["attn", "gemm", "rows"]
    if "attn" in what:
        bench_attn()
        bench_attn(B=1)
        bench_attn(Lk=512)
        bench_attn(Lk=257)
        bench_attn_dual()
    if "attnself" in what:
        bench_attn()
        bench_attn(B=1, Lq=28800, Lk=28800, nbuf=2)
    if "attncross" in what:
        bench_attn(Lk=512)
        bench_attn(Lk=257)
        bench_attn_dual()
    if "attnlib" in what:
        bench_attn()
        bench_attn_libs()
        bench_attn(B=1, Lq=28800, Lk=28800, nbuf=2)
        bench_attn_libs(B=1, L=28800)
        bench_attn(Lk=512)
        bench_attn_libs(L=7200, Lk=512)
    if "gemm" in what:
        for (M, N, K, epi) in [(14400, 15360, 5120, 0), (14400, 5120, 5120, 3), (14400, 13824, 5120, 1), (14400, 5120, 13824, 3), (7200, 5120, 5120, 0)]:
            bench_gemm(M, N, K, epi)
    if "rows" in what:
        bench_rows()
    if "conv" in what:
        bench_conv(96, 96, 4, 720, 1280)
        bench_conv(192, 192, 4, 360, 640)
        bench_conv(384, 384, 2, 180, 320)
