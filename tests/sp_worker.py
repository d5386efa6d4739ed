"""Worker of tests/test_gpu_seqpar.py (run under torchrun, one process per GPU): the sequence-parallel forward must be bit-identical
to the single-GPU forward of the same model on every rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import chronoedit_b200 as ce
    from chronoedit_b200 import parallel

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # (heads, dim = heads*128, latent T, H, W): 512 tokens (256 per rank at world 2) and 1920 tokens (960 per rank: query tiles of 128 rows
    # straddle the rank boundary), batch 2
    for heads, (T, H, W), B in ((2, (2, 32, 32), 1), (4, (2, 48, 40), 2), (8, (8, 24, 20), 1)):
        if heads % world or (T * (H // 2) * (W // 2)) % world:
            continue
        m = ce.ChronoEditTransformer3DModel(num_attention_heads=heads, in_channels=36, out_channels=16, ffn_dim=1024, num_layers=2, image_dim=1280,
                                            added_kv_proj_dim=heads * 128, device=dev)
        g = torch.Generator(device=dev).manual_seed(1234)          # same seed on every rank: identical weights
        for n, p in m.named_parameters():
            if ".norm" in n and n.endswith("weight"):
                p.data.normal_(0, 0.1, generator=g).add_(1.0)
            else:
                p.data.normal_(0, 0.03, generator=g)
        gi = torch.Generator(device=dev).manual_seed(99)
        x = torch.randn(B, 36, T, H, W, generator=gi, device=dev).bfloat16()
        text = torch.randn(B, 512, 4096, generator=gi, device=dev).bfloat16()
        img = torch.randn(B, 257, 1280, generator=gi, device=dev).bfloat16()
        t = torch.full((B,), 613, device=dev)
        single = m(x, t, text, img, return_dict=False)[0].clone()
        parallel.enable_sequence_parallel(m, B, T, H, W)
        for it in range(3):
            out = m(x, t, text, img, return_dict=False)[0]
            torch.cuda.synchronize()
            same = torch.equal(out, single)
            print(f"rank {rank}: heads={heads} latent=({T},{H},{W}) B={B} iter {it}: sequence-parallel == single-GPU: {same}"
                  + ("" if same else f"  max|diff|={float((out.float() - single.float()).abs().max()):.4g}"), flush=True)
            ok = ok and same
        parallel.disable_sequence_parallel(m)
        again = m(x, t, text, img, return_dict=False)[0]
        ok = ok and torch.equal(again, single)
        del m
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
