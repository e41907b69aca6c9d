"""Summarise an ncu launch list (gpu__time_duration.sum per launch) into per-kernel totals and shares.
usage: python tools/launch_shares.py gpurun_out/launches.csv profiles/summary.txt [title]"""
import csv, sys
from collections import defaultdict
src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = r[ki]
    name = name.split("(")[0].replace("void ", "").replace("rmu::", "")
    if not name.startswith(("scan_", "finalize", "exact", "select_tau", "compact", "merge", "gemm_", "attention", "ln_kernel",
                            "embed_ln", "pool_", "cls_head", "row_stats", "gather_rows", "mmr", "split_planes", "join_planes", "bm25_")):
        name = "torch/other: " + name[:40]
    tot[name] += v * scale; cnt[name] += 1
ours = {k: v for k, v in tot.items() if not k.startswith("torch/other")}
s_ours = sum(ours.values())
with open(out, "w") as f:
    f.write(f"# {title}\n# source: {src}  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
    f.write(f"# launches of this library: {sum(cnt[k] for k in ours)}; device time {s_ours/1e3:.2f} ms; other (torch) {sum(tot.values())/1e3 - s_ours/1e3:.2f} ms\n")
    f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share_of_ours':>14s}\n")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
        f.write(f"{k[:70]:70s} {cnt[k]:8d} {v:12.1f} {v/cnt[k]:10.1f} {100*v/s_ours:13.1f}%\n")
    other = sorted(((k, v) for k, v in tot.items() if k.startswith("torch/other")), key=lambda kv: -kv[1])[:8]
    for k, v in other:
        f.write(f"{k[:70]:70s} {cnt[k]:8d} {v:12.1f} {v/cnt[k]:10.1f}\n")
print(open(out).read())
