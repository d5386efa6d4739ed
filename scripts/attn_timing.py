#!/usr/bin/env python
"""Phase timing of one softmax warp (block 0) for the attention kernels: python scripts/attn_timing.py"""
import math, os, sys, torch
sys.path.insert(0, ".")
import chronoedit_b200._lib as L
lib = L.lib()
B, H, Lq = 1, 40, 7200
D = H * 128
x = torch.randn(B, Lq, 3 * D, device="cuda", dtype=torch.bfloat16)
out = torch.empty(B, Lq, D, device="cuda", dtype=torch.bfloat16)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
L.check(lib.ce_debug_attention_timing(L.ptr(buf)))
for _ in range(3):
    L.check(lib.ce_attention_bf16(L.ptr(x), 3 * D, L.ptr(x[..., D:]), 3 * D, L.ptr(x[..., 2 * D:]), 3 * D, L.ptr(out), D, B, H, Lq, Lq, 1 / math.sqrt(128), 0,
                                  L.current_stream()))
torch.cuda.synchronize()
t = buf.cpu().tolist()
ver = os.environ.get("CE_ATTN_V2", "2")
if ver == "6":
    # attention6.cu: event log only (block 0, key tiles 16..23): per query tile and half-row thread
    t0 = t[64]
    print("kernel: attention6 (two softmax threads per row); cycles relative to 'S(16) seen' by tile 0 / half 0")
    for qt in range(2):
        for hf in range(2):
            print(f"query tile {qt}, half {hf}:  j | S seen | S in regs | max exchanged | first P part | last P part" + (" || first P.V issued | last P.V issued | S(j+1) issued" if hf == 0 else ""))
            for k in range(8):
                e = t[64 + qt * 128 + hf * 64 + 8 * k: 64 + qt * 128 + hf * 64 + 8 * k + 8]
                line = "  j=%d  %7d %7d %7d %7d %7d" % (16 + k, e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0)
                if hf == 0:
                    line += " || %7d %7d %7d" % (e[5] - t0, e[6] - t0, e[7] - t0)
                print(line)
    sys.exit(0)
v2 = ver != "0"
names = (["wait S", "ld S + free buffer", "max", "wait m(j-1), decide, publish m(j)", "(rescale)", "wait exp turn, exp, wait P.V(j-2)", "store P, arrive"] if ver == "5" else
         ["wait S", "ld S", "max", "rescale, exp of keys 0-79, publish P[0:64)", "exp of keys 80-127, publish P[64:128)"]) if v2 else ["wait S", "ld S", "max+decide", "exp+pack", "wait PV(t-1)/rescale", "store P + arrive"]
n = t[7] if ver == "5" else (t[5] if v2 else t[6])
print("kernel:", ("attention5 (cta_group::2)" if ver == "5" else "attention2 (v2)") if v2 else "attention (v1)", "tiles of this group:", n)
for nm, c in zip(names, t):
    print(f"  {nm:28s} {c / max(n,1):9.1f} cycles/tile")
print(f"  {'total':28s} {sum(t[:len(names)]) / max(n,1):9.1f} cycles/tile")
if ver == "5" and t[15]:
    print("peer CTA of the same cluster:")
    for nm, c in zip(names, t[8:15]):
        print(f"  {nm:28s} {c / max(t[15],1):9.1f} cycles/tile")
    print(f"  {'total':28s} {sum(t[8:15]) / max(t[15],1):9.1f} cycles/tile")

if v2 and ver != "5" and any(t[64:192]):
    # event log of block 0, key tiles 16..23 (SM clock), both query tiles: softmax thread: S ready seen, P[0:64) published,
    # P[64:128) published; issuer thread: P.V first half issued, second half issued, S(j) issued (the S that tile j consumes)
    t0 = t[64]
    for qt in range(2):
        print(f"event log, query tile {qt} (cycles relative to 'S(16) seen' of query tile 0): tile | S seen | P1 pub | P2 pub || PV1 issued | PV2 issued | S(j) issued")
        for k in range(8):
            e = t[64 + 64 * qt + 8 * k: 64 + 64 * qt + 8 * k + 6]
            print("  j=%d  %7d %7d %7d || %7d %7d %7d" % (16 + k, e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0))
