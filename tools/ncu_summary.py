"""Summarise an .ncu-rep (raw page) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_scan.txt [title]"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
with open(out, "w") as f:
    f.write(f"# {title}\n# source: {rep} (ncu --set full --clock-control none)\n")
    for r in rows[2:]:
        f.write("\n")
        for h, u, v in zip(hdr, units, r):
            key = h.split("TriageCompute.")[-1]
            if key in WANT or h in WANT:
                f.write(f"{key:90s} {u:16s} {v}\n")
print(open(out).read())
