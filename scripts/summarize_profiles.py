#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
    python scripts/summarize_profiles.py r01
"""
import csv, json, os, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)
# 1. launch list: per-kernel total time / count / share
path = f"gpurun_out/launches_{tag}.csv"
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
    with open(f"profiles/{tag}_launches_summary.csv", "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{ns/1e3:.1f},{ns/tot:.4f}\n")
    print(open(f"profiles/{tag}_launches_summary.csv").read())
# 2. full captures: selected metrics per captured launch
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_tensor.sum"]
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
