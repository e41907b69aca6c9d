"""Summarise an .ncu-rep for one kernel: duration, issue utilisation, stall reasons, opcode mix, hottest instructions.
usage: python tools/ncu_summary.py report.ncu-rep [n_hot]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
print("kernel", d.get("Kernel Name", "")[:80])
for k in ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"):
    if k in d: print(f"  {k} = {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); ops = collections.Counter(); total = 0; data = []
for r in rows[2:]:
    if not r[ix["# Samples"]].isdigit(): continue
    n = int(r[ix["# Samples"]]); total += n; data.append(r)
    for s in stalls:
        if r[ix[s]].isdigit(): tot[s] += int(r[ix[s]])
    w = r[ix["Source"]].split()
    if w:
        o = w[1] if w[0].startswith("@") and len(w) > 1 else w[0]
        ops[o.split(".")[0]] += int(r[ix["Instructions Executed"]]) if r[ix["Instructions Executed"]].isdigit() else 0
print("stall samples", total, {s[6:]: round(100 * v / max(total, 1), 1) for s, v in tot.most_common(9)})
it = sum(ops.values())
print("opcode mix (% of executed warp instructions):", {k: round(100 * v / max(it, 1), 1) for k, v in ops.most_common(18)})
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:nhot]:
    ss = sorted(((s[6:], int(r[ix[s]])) for s in stalls if r[ix[s]].isdigit() and int(r[ix[s]]) > 0), key=lambda x: -x[1])[:2]
    print(f"  {r[ix['Address']][-5:]} samples {r[ix['# Samples']]:>5} exec {r[ix['Instructions Executed']]:>9}  {r[ix['Source']][:64]:64s} {ss}")
