#!/usr/bin/env python
"""Benchmark of the ChronoEdit denoising hot path (BASELINE.json metric: DiT denoising steps/sec at 14B, 720x1280,
5 pixel frames = 2 latent frames).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one iteration of the sampling loop of ChronoEditPipeline.__call__ (pipeline_chronoedit.py:695-753) for ONE
edit with classifier-free guidance: two DiT forwards (prompt / negative prompt — evaluated as one batch-2 call, which the
batch-invariance test proves equal to two batch-1 calls), the CFG combine and the latent update.  Workload = configs[1]:
ChronoEdit-14B (40 layers, dim 5120, ffn 13824), latent [1,36,2,90,160] -> 7200 tokens, 512 text + 257 image tokens,
random-init bf16 weights, synthetic inputs.  With N GPUs every rank runs its own independent edit (data parallel over a
batch of N edits, weak scaling); rank 0 initialises the weights and they are replicated with ONE NCCL broadcast; there
is no per-step collective.

Printed JSON (one line, rank 0):
  value     steps/s summed over ranks, inputs resident in HBM, timed with CUDA events, max over ranks
  e2e       the same step through the host-buffer C-ABI call (ce_dit_forward_host): pinned-host inputs H2D every step,
            sample D2H every step, CFG combine + latent update on the host
  roofline  the dominant kernel class (tcgen05 GEMM): algorithmic FLOPs / device time of those launches measured live
            with CUDA events inside the timed region, against MEASURED_PEAKS.json bf16_tflops_sustained
  cpu_baseline  the oracle (CPU restatement of the reference) timed on the host cores on a bounded sample
--impl reference times that CPU path alone with the same metric/config (the reference has no faster path on this box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dit_denoising_steps_per_sec_14B_720p_5frame"
FRAMES, LAT_H, LAT_W, TEXT_LEN = 2, 90, 160, 512
GUIDANCE = 5.0


def model_config(layers: int):
    return dict(patch_size=(1, 2, 2), num_attention_heads=40, attention_head_dim=128, in_channels=36, out_channels=16,
                text_dim=4096, freq_dim=256, ffn_dim=13824, num_layers=layers, image_dim=1280, added_kv_proj_dim=5120)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------
# CPU side: the oracle on a bounded sample (cpu_baseline of our arm, and the whole of --impl reference)
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(reps: int, warmup: int = 0):
    """Times ONE ChronoEditTransformerBlock at 14B width (dim 5120, 40 heads, ffn 13824, 769 context tokens) on a
    quarter of the tokens (1800 = 1 frame x 45 x 40 patches) in fp32 on all host cores with the oracle, `reps` times.
    A full step is 2 forwards x 40 blocks at 7200 tokens; the step time is extrapolated by the algorithmic FLOP ratio
    (attention's quadratic term included).  Returns (steps_per_sec, seconds_per_sample list, cores, description)."""
    import torch

    from oracle import dit_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.DiTConfig.chronoedit_14b()
    D = cfg.inner_dim
    one = O.DiTConfig(num_layers=1)
    g = torch.Generator().manual_seed(0)
    sd = {k: v for k, v in O.random_state_dict(one, seed=0).items() if k.startswith("blocks.0.")}
    Ls = 1800
    x = torch.randn(1, Ls, D, generator=g)
    ctx = torch.randn(1, 257 + 512, D, generator=g)
    temb6 = torch.randn(1, 6, D, generator=g) * 0.1
    freqs = O.rope_table(cfg, 2, 90, 160)[:, :, :Ls]
    times = []
    with torch.no_grad():
        # give the reference its best thread count (oversubscribed SMT threads often hurt torch's CPU GEMMs)
        best = None
        for n in sorted({cores, max(1, cores // 2), max(1, cores // 4)}, reverse=True):
            torch.set_num_threads(n)
            t0 = time.perf_counter()
            O.block(sd, 0, one, x, ctx, temb6, freqs)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, n)
        cores = best[1]
        torch.set_num_threads(cores)
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            O.block(sd, 0, one, x, ctx, temb6, freqs)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    f_sample = O.flops_per_forward(O.DiTConfig(num_layers=1), 2, 90, 160) - O.flops_per_forward(O.DiTConfig(num_layers=0), 2, 90, 160)
    # FLOPs of one block at Ls tokens (same formula, linear + quadratic terms)
    Fd = cfg.ffn_dim
    def blk(L):
        return (2 * L * D * 3 * D + 4 * L * L * D + 2 * L * D * D + 2 * L * D * D + 2 * 769 * D * 2 * D + 4 * L * 769 * D
                + 2 * L * D * D + 4 * L * D * Fd)
    scale = blk(7200) / blk(Ls)
    t_block = statistics.mean(times) * scale
    step_s = 2 * 40 * t_block
    desc = (f"one 14B-width DiT block (oracle, fp32) on {Ls} of 7200 tokens x {reps} reps, {cores} threads; step time = sample x "
            f"{scale:.2f} (FLOP ratio) x 40 blocks x 2 CFG forwards [extrapolated]")
    assert abs(blk(7200) - f_sample) / f_sample < 1e-6
    return 1.0 / step_s, times, cores, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    rate, times, cores, desc = cpu_reference_step_rate(args.steps, args.warmup)
    ms_per_step = 1000.0 / rate
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ChronoEdit-14B DiT, 720x1280, 5 px frames (7200 tokens), CFG step = 2 forwards", "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 2),
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
# Algorithmic work, restated here so that the measured arm does not touch oracle/ at all:
# 2*MAC for every Linear, 4*Lq*Lk*D for every attention; tests/test_oracle.py checks these against the oracle's own counters.
def dit_flops_per_forward(layers: int, frames: int, lat_h: int, lat_w: int, text_len: int, image_len: int, batch: int) -> float:
    D, Fd, cin, cout, text_dim, image_dim, freq_dim = 5120, 13824, 36, 16, 4096, 1280, 256
    L = frames * (lat_h // 2) * (lat_w // 2)
    per_block = (2 * L * D * 3 * D + 4 * L * L * D + 2 * L * D * D            # self: q,k,v / SDPA / out
                 + 2 * L * D * D + 2 * text_len * D * 2 * D + 2 * image_len * D * 2 * D   # cross: q / text k,v / image k,v
                 + 4 * L * (text_len + image_len) * D + 2 * L * D * D          # cross: SDPA / out
                 + 2 * 2 * L * D * Fd)                                          # ffn
    embed = (2 * L * cin * 4 * D + 2 * L * D * cout * 4 + 2 * text_len * (text_dim * D + D * D)
             + 2 * image_len * (image_dim ** 2 + image_dim * D) + 2 * (freq_dim * D + D * D + D * 6 * D))
    return float(batch) * (layers * per_block + embed)


VAE_ENCODE_FLOP = 24577494220800.0   # conv FLOPs of one 5 x 720 x 1280 encode / decode (24.58 / 41.04 TFLOP)
VAE_DECODE_FLOP = 41036724633600.0


def init_weights_(model, seed: int):
    """Random-init weights of the 14B architecture directly on the device (there is no checkpoint on the box):
    Linear ~ N(0, 0.02), norms ~ 1 + 0.1 N, scale_shift_table ~ N(0,1)/sqrt(D) (as transformer_chronoedit.py:265, 393)."""
    import math

    import torch

    g = torch.Generator(device=model.device).manual_seed(seed)
    D = model.config.num_attention_heads * model.config.attention_head_dim
    for n, p in model.named_parameters():
        if n.endswith("scale_shift_table"):
            p.data.normal_(0, 1.0 / math.sqrt(D), generator=g)
        elif ".norm" in n and n.endswith("weight"):
            p.data.normal_(0, 0.1, generator=g).add_(1.0)
        elif ".norm" in n and n.endswith("bias"):
            p.data.normal_(0, 0.1, generator=g)
        else:
            p.data.normal_(0, 0.02, generator=g)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import chronoedit_b200 as ce
    from chronoedit_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = model_config(args.layers)
    model = ce.ChronoEditTransformer3DModel(**cfg, device=dev)
    # weights: rank 0 initialises, ONE broadcast replicates them (the only collective on the path)
    t_b0 = time.perf_counter()
    if rank == 0:
        init_weights_(model, seed=0)
    model.pack_weights()
    bcast_bytes = 0
    if world > 1:
        from chronoedit_b200 import parallel

        bcast_bytes = parallel.broadcast_module_weights(model, src=0)
        torch.cuda.synchronize()
    t_bcast = time.perf_counter() - t_b0

    # synthetic per-edit inputs (different on every rank: independent edits)
    g = torch.Generator(device="cpu").manual_seed(42 + rank)
    B = 2  # CFG pair
    latents = torch.randn(1, 16, FRAMES, LAT_H, LAT_W, generator=g)
    cond = torch.randn(1, 20, FRAMES, LAT_H, LAT_W, generator=g)
    cond[:, :4] = 0
    cond[:, :4, 0] = 1  # mask channels: frame 0 = 1 (pipeline_chronoedit.py:447-453)
    text = torch.randn(2, TEXT_LEN, 4096, generator=g)
    text[0, 120:] = 0
    text[1, 40:] = 0
    img = torch.randn(1, 257, 1280, generator=g).expand(2, -1, -1).contiguous()
    sigmas = torch.linspace(1.0, 0.0, 51)
    sigmas = 5.0 * sigmas / (1 + 4.0 * sigmas)  # flow shift 5 (run_inference_diffusers.py:203-207)

    d_lat = latents.to(dev)
    d_cond = cond.to(dev, torch.bfloat16)
    d_text = text.to(dev, torch.bfloat16)
    d_img = img.to(dev, torch.bfloat16)

    # The loop body of ChronoEditPipeline.__call__ (pipeline_chronoedit.py:693-739) on device-resident tensors: model input =
    # cat([latents, condition]) in bf16, ONE batch-2 DiT call for the CFG pair, then ONE fused launch for the guidance combine
    # + UniPC flow-matching scheduler step + the latent channels of the next model input (chronoedit_b200/scheduler.py).
    from chronoedit_b200.scheduler import FlowUniPCMultistepScheduler

    sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    d_lat = d_lat.to(torch.bfloat16)   # the diffusers pipeline keeps bf16 latents (pipeline_chronoedit.py:676-687)
    d_in = torch.cat([d_lat, d_cond], dim=1).contiguous()

    def device_step(i):
        nonlocal d_lat
        if sched.step_index is None or sched.step_index >= 50:
            sched.set_timesteps(50, device=dev, shift=5.0)   # flow shift 5 (run_inference_diffusers.py:203-207)
        t = sched.timesteps[sched.step_index or 0]
        out = model(d_in.expand(2, -1, -1, -1, -1), t.expand(2), d_text, d_img, return_dict=False)[0]
        d_lat = sched.step_cfg(out[0:1], out[1:2], GUIDANCE, t, d_lat, model_input_out=d_in)

    # host-buffer (e2e) step: pinned inputs, H2D + forward + D2H inside the C-ABI call, glue on the host
    h_x = torch.empty(2, 36, FRAMES, LAT_H, LAT_W, dtype=torch.bfloat16).pin_memory()
    h_text = text.to(torch.bfloat16).pin_memory()
    h_img = img.to(torch.bfloat16).pin_memory()
    h_out = torch.empty(2, 16, FRAMES, LAT_H, LAT_W, dtype=torch.bfloat16).pin_memory()
    h_lat = latents.clone()
    h_cond = cond.to(torch.bfloat16)
    h2d = h_x.numel() * 2 + h_text.numel() * 2 + h_img.numel() * 2 + B * 4
    d2h = h_out.numel() * 2

    def host_step(i):
        nonlocal h_lat
        s0, s1 = float(sigmas[i % 50]), float(sigmas[i % 50 + 1])
        t = torch.full((B,), float(int(s0 * 1000) % 1000), dtype=torch.float32)
        xi = torch.cat([h_lat.to(torch.bfloat16), h_cond], dim=1)
        h_x[0].copy_(xi[0]); h_x[1].copy_(xi[0])
        model.forward_host(h_x, t, h_text, h_img, out=h_out)
        o = h_out.float()
        h_lat = h_lat + (s1 - s0) * (o[1:2] + GUIDANCE * (o[0:1] - o[1:2]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, profile=False):
        for i in range(warmup):
            fn(i)
        barrier()
        if sampler:
            sampler.start()
        if profile:
            _lib.check(_lib.lib().ce_dit_profile_begin(model._handle, 1200 * steps))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        prof = None
        if profile:
            import ctypes
            ms = (ctypes.c_double * 4)(); work = (ctypes.c_double * 4)(); cnt = (ctypes.c_int64 * 4)()
            _lib.check(_lib.lib().ce_dit_profile_end(model._handle, ms, work, cnt))
            prof = {"ms": list(ms), "work": list(work), "count": list(cnt)}
        clocks = sampler.stop() if sampler else None
        barrier()
        dev_ms = e0.elapsed_time(e1)
        t = torch.tensor([dev_ms, wall * 1000.0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), prof, clocks

    sampler = ClockSampler(local) if rank == 0 else None
    dev_ms, wall_ms, prof, clocks = timed(device_step, args.steps, args.warmup, sampler, profile=True)
    launches = (model.launches_per_forward() + 1) * args.steps   # + the fused sampler launch
    value = world * args.steps / (dev_ms / 1000.0)
    e2e_dev_ms, e2e_wall_ms, _, _ = timed(host_step, max(3, args.steps // 2), 1)
    e2e_steps = max(3, args.steps // 2)
    e2e_value = world * e2e_steps / (e2e_wall_ms / 1000.0)

    # the VAE bookends of one edit (encode of the condition video, decode of the result), timed once per run
    edit = None
    if not args.no_vae:
        vae = ce.AutoencoderKLWan(device=dev)
        gv = torch.Generator(device=dev).manual_seed(7)
        for n, p in vae.named_parameters():
            if n.endswith("gamma"):
                p.data.normal_(0, 0.1, generator=gv).add_(1.0)
            elif n.endswith("bias"):
                p.data.normal_(0, 0.02, generator=gv)
            else:
                fan_in = p[0].numel()
                p.data.normal_(0, 1.0 / fan_in ** 0.5, generator=gv)
        video = torch.zeros(1, 3, 5, 8 * LAT_H, 8 * LAT_W, dtype=torch.bfloat16, device=dev)
        video[:, :, 0] = torch.rand(1, 3, 8 * LAT_H, 8 * LAT_W, device=dev) * 2 - 1
        zlat = torch.randn(1, 16, FRAMES, LAT_H, LAT_W, dtype=torch.bfloat16, device=dev)

        def time_once(fn, reps=2):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        enc_ms = time_once(lambda: vae.encode(video))
        enc_launches = vae.launches()
        dec_ms = time_once(lambda: vae.decode(zlat))
        dec_launches = vae.launches()
        step_ms = dev_ms / args.steps
        edit = {
            "vae_encode_ms": enc_ms, "vae_decode_ms": dec_ms, "vae_encode_launches": enc_launches, "vae_decode_launches": dec_launches,
            "vae_encode_conv_tflops": VAE_ENCODE_FLOP / enc_ms / 1e9,
            "vae_decode_conv_tflops": VAE_DECODE_FLOP / dec_ms / 1e9,
            "vae_decode_algorithmic_GBps": 23.64e9 / dec_ms / 1e6,
            "edits_per_sec_50_steps_all_gpus": world / ((enc_ms + dec_ms + 50 * step_ms) / 1000.0),
            "edits_per_sec_8_steps_no_cfg_all_gpus": world / ((enc_ms + dec_ms + 8 * step_ms / 2) / 1000.0),
            "note": "edit = VAE encode + N denoising steps + VAE decode (SURVEY 8d); text/image encoders excluded ('next' row)",
        }
        del vae
        torch.cuda.empty_cache()

    if rank == 0:
        peak_tf, peak_hbm, peak_src = peaks()
        flops_fwd = dit_flops_per_forward(args.layers, FRAMES, LAT_H, LAT_W, TEXT_LEN, 257, batch=2)
        gemm_tf = prof["work"][0] / (prof["ms"][0] / 1000.0) / 1e12 if prof["ms"][0] > 0 else 0.0
        attn_tf = prof["work"][1] / (prof["ms"][1] / 1000.0) / 1e12 if prof["ms"][1] > 0 else 0.0
        kernel_ms = sum(prof["ms"])
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": ("configs[1]: ChronoEdit-14B single edit, 720x1280, 5 px frames -> latent [1,36,2,90,160] (7200 tokens), "
                             "512 text + 257 image tokens, CFG 5.0 (2 forwards/step as one batch-2 call), per-GPU independent edits")
                if FRAMES == 2 else
                ("DEV (not the headline workload): configs[2] temporal-reasoning geometry, 29 px frames -> latent [1,36,8,90,160] "
                 "(28800 tokens), CFG 5.0"),
                "layers": args.layers, "global_batch_edits": world, "parallelism": f"dp{world}",
                "l2": "inputs larger than L2 (32.8 GB of weights stream every forward); no explicit flush",
                "latent_update": "value: fused CFG + FlowUniPC step + next model input in one launch (ce_unipc_step); e2e: host-side Euler glue around ce_dit_forward_host",
                "cross_kv_hoisting": False,
            },
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_wall_ms / e2e_steps},
            "gpu_launches": launches,
            "roofline": {
                "bound": "tensor", "kernel": "gemm_bf16_2cta_kernel / gemm_bf16_kernel (tcgen05, all Linear layers)",
                "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf, "peak_source": peak_src + " bf16_tflops_sustained",
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at the QKV shape (M=14400, N=15360, K=5120)
                # from the committed ncu --set full capture profiles/r01w_ncu_gemm.csv; algorithmic operand bytes of that launch
                # = (M*K + N*K + M*N)*2 = 0.747e9 (each L2 die fetches its own copy of A and W)
                "traffic": 1.401254e9 + 0.430177e9, "traffic_launch": "gemm_bf16_2cta_kernel M=14400 N=15360 K=5120 (profiles/r01w_ncu_gemm.csv)",
                "algorithmic_bytes_of_that_launch": (14400 * 5120 + 15360 * 5120 + 14400 * 15360) * 2,
                "launches": prof["count"][0], "ms_total": prof["ms"][0],
                "share_of_kernel_time": prof["ms"][0] / kernel_ms if kernel_ms else None,
                "attention": {"achieved": attn_tf, "frac": attn_tf / peak_tf, "ms_total": prof["ms"][1], "launches": prof["count"][1]},
                "rows_ms_total": prof["ms"][2], "other_ms_total": prof["ms"][3],
                "whole_step": {"algorithmic_tflop_per_step": flops_fwd / 1e12, "achieved": flops_fwd / 1e12 / (dev_ms / args.steps / 1000.0),
                               "frac": flops_fwd / 1e12 / (dev_ms / args.steps / 1000.0) / peak_tf},
            },
            "clocks": clocks,
            "weight_broadcast": {"bytes": bcast_bytes, "seconds_incl_init": round(t_bcast, 3)},
            "edit": edit,
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, times, cores, desc = cpu_reference_step_rate(reps=3, warmup=1)
            line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=40, help="DEV ONLY: fewer layers make the number invalid as a bench value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vae", action="store_true", help="skip the (untimed-region) VAE encode/decode measurement")
    ap.add_argument("--latent-frames", type=int, default=2, choices=[2, 8],
                    help="DEV ONLY: 8 = the temporal-reasoning geometry of configs[2] (28 800 tokens); not the headline workload")
    args = ap.parse_args()
    if args.latent_frames != 2:
        global FRAMES
        FRAMES = args.latent_frames
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
