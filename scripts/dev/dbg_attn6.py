import sys, math, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chronoedit_b200 import _lib as L
lib = L.lib()
L.check(lib.ce_debug_attention_kernel(6))
B, H, Lq, Lk, hd = 1, 2, 256, 384, 128
D = H * hd
g = torch.Generator().manual_seed(1)
q = torch.randn(B, Lq, D, generator=g).bfloat16().cuda(); k = torch.randn(B, Lk, D, generator=g).bfloat16().cuda(); v = torch.randn(B, Lk, D, generator=g).bfloat16().cuda()
out = torch.zeros(B, Lq, D, dtype=torch.bfloat16, device='cuda')
L.check(lib.ce_attention_bf16(L.ptr(q), D, L.ptr(k), D, L.ptr(v), D, L.ptr(out), D, B, H, Lq, Lk, 1.0 / math.sqrt(hd), 0, L.current_stream()))
torch.cuda.synchronize()
ref = torch.nn.functional.scaled_dot_product_attention(q.view(B, Lq, H, hd).transpose(1, 2).float(), k.view(B, Lk, H, hd).transpose(1, 2).float(), v.view(B, Lk, H, hd).transpose(1, 2).float()).transpose(1, 2).reshape(B, Lq, D)
d = (out.float() - ref).abs()
print("max err", d.max().item(), "mean", d.mean().item())
print("per 64-col block max err:", [round(d[..., i*64:(i+1)*64].max().item(), 4) for i in range(D // 64)])
print("rows blocks:", [round(d[0, i*32:(i+1)*32].max().item(), 4) for i in range(Lq // 32)])
