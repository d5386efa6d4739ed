#!/usr/bin/env python
"""Single-edit latency with sequence parallelism: ONE edit's tokens split over the GPUs of a node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 scripts/bench_seqpar.py [--latent-frames 8] [--layers 40]

Workload = BASELINE configs[2] geometry by default: ChronoEdit-14B, 8 latent frames at 720p = 28 800 tokens, CFG pair as one batch-2
forward per step.  Prints ms per step with the tokens split over W ranks next to the single-GPU forward measured in the same run
(rank 0, sequence parallelism off), and checks the two outputs are bit-identical."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    import chronoedit_b200 as ce
    from chronoedit_b200 import parallel

    ap = argparse.ArgumentParser()
    ap.add_argument("--latent-frames", type=int, default=8, choices=[2, 8])
    ap.add_argument("--layers", type=int, default=40)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model = ce.ChronoEditTransformer3DModel(**bench.model_config(args.layers), device=dev, cache_context=True)
    bench.init_weights_(model, seed=0)          # same seed on every rank: identical weights without a broadcast
    model.pack_weights()
    T, H, W = args.latent_frames, 90, 160
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(2, 36, T, H, W, generator=g, device=dev).bfloat16()
    text = torch.randn(2, 512, 4096, generator=g, device=dev).bfloat16()
    img = torch.randn(2, 257, 1280, generator=g, device=dev).bfloat16()
    t = torch.full((2,), 500, device=dev)

    def timed(n):
        model(x, t, text, img, return_dict=False)     # warm-up (context cache, workspaces)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = model(x, t, text, img, return_dict=False)[0]
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out

    single_ms, single = timed(args.steps)              # every rank runs the whole edit by itself
    region = parallel.enable_sequence_parallel(model, 2, T, H, W)
    model.clear_context_cache()
    sp_ms, out = timed(args.steps)
    same = torch.equal(out, single)
    flag = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"workload": f"ChronoEdit-14B ({args.layers} layers), latent [2,36,{T},90,160] = {T * 45 * 80} tokens x CFG pair, one step",
                          "world": world, "ms_per_step_single_gpu": single_ms, "ms_per_step_sequence_parallel": sp_ms,
                          "speedup": single_ms / sp_ms, "efficiency": single_ms / sp_ms / world, "bit_identical_on_all_ranks": bool(flag.item()),
                          "peer_region_bytes_per_rank": region}), flush=True)
    parallel.disable_sequence_parallel(model)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
