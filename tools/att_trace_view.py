"""Render gpurun_out/att_trace.txt (RMU_ATTN_TRACE=1: event, arg, 32-bit clock of CTA 0) as a per-tile timeline."""
import sys
ev_names = {1: "kv load issued (work)", 2: "q load issued (tile)", 3: "QK: q_full seen (tile)", 4: "QK g0 issued (tile)", 5: "QK g1 issued (tile)",
            6: "QK g0 MMAs+commit issued", 7: "QK g1 MMAs+commit issued", 12: "PV g0 chunk MMAs issued", 13: "PV g1 chunk MMAs issued",
            10: "PV g0 chunk seen", 11: "PV g1 chunk seen", 20: "SM g0 s_full (tile)", 21: "SM g1 s_full (tile)", 22: "SM g0 pass1+xchg done",
            23: "SM g1 pass1+xchg done", 24: "SM g0 chunk published", 25: "SM g1 chunk published", 26: "OUT g0 o_full seen", 27: "OUT g1 o_full seen",
            28: "OUT g0 done", 29: "OUT g1 done"}
rows = [tuple(map(int, l.split())) for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/att_trace.txt")]
rows.sort(key=lambda r: r[2])
t0 = rows[0][2]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 160)
for ev, arg, t in rows[lo:hi]:
    print(f"{t - t0:9d}  {ev_names.get(ev, ev):28s} {arg}")
