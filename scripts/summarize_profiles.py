#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
    python scripts/summarize_profiles.py r01
"""
import csv, json, os, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)
# 1. launch list: per-kernel total time / count / share
for path, outp in ((f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launches_summary.csv"), (f"gpurun_out/launches_vae_{tag}.csv", f"profiles/{tag}_launches_vae_summary.csv")):
  if os.path.exists(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi: continue
        name = r[ki].split("(")[0].replace("void ", "")[-70:]
        v = float(r[vi].replace(",", "")); u = r[ui]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    with open(outp, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{ns/1e3:.1f},{ns/tot:.4f}\n")
    print(open(outp).read())
# 2. full captures: selected metrics per captured launch
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_tensor.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for k in ("gemm", "attn", "xattn", "rows", "conv"):
    rep = f"gpurun_out/prof_{k}_{tag}.ncu-rep"
    if not os.path.exists(rep): continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(f"profiles/{tag}_ncu_{k}.csv", "w") as f:
        cols = ["Kernel Name", "Grid Size", "Block Size"] + [w for w in want if w in hdr]
        f.write(",".join(cols) + "\n")
        f.write(",".join(units[hdr.index(c)] for c in cols) + "\n")
        for r in rows[2:]:
            f.write(",".join('"' + r[hdr.index(c)].replace('"', "'")[:90] + '"' for c in cols) + "\n")
    print(f"profiles/{tag}_ncu_{k}.csv written ({len(rows)-2} launches)")

# 3. profiles/traffic.json: per-launch DRAM traffic of the dominant kernel classes (what bench.py's roofline.*.traffic reads),
#    taken from the newest profiles/*_ncu_<class>.csv that exists; algorithmic bytes = the launch's operands touched once.
def _first_row(path):
    rows = list(csv.reader(open(path)))
    hdr, units, r = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
    return r[0], rd, wr, float(r[hdr.index("gpu__time_duration.sum")])

def _newest(kind):
    import glob
    c = sorted(glob.glob(f"profiles/*_ncu_{kind}.csv"), key=os.path.getmtime)
    return c[-1] if c else None

LAUNCH = {  # the launch each capture script profiles (scripts/gpu_profile.sh -> scripts/bench_ops.py) and its algorithmic bytes
    "gemm": ("gemm", "QKV GEMM M=14400 N=15360 K=5120 (bf16 A, W, out)", (14400 * 5120 + 15360 * 5120 + 14400 * 15360) * 2),
    "attention": ("attn", "self-attention B=2 H=40 L=7200 hd=128 (q, k, v, out once)", 4 * 2 * 7200 * 5120 * 2),
    "conv": ("conv", "conv 3x3x3 96->96 on 4x720x1280 (input 6 frames incl. causal history + output 4 frames, bf16)", (6 + 4) * 720 * 1280 * 96 * 2),
}
traffic = {}
for name, (kind, desc, alg) in LAUNCH.items():
    pth = _newest(kind)
    if not pth:
        continue
    kname, rd, wr, ms = _first_row(pth)
    traffic[name] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": kname[:80], "launch": f"{desc} ({pth})",
                     "algorithmic_bytes": alg, "ratio": round((rd + wr) / alg, 3), "ncu_ms": ms}
with open("profiles/traffic.json", "w") as f:
    json.dump(traffic, f, indent=1)
print(json.dumps(traffic, indent=1))
