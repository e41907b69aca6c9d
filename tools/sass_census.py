"""SASS instruction census of the built library: which Blackwell instructions each hot kernel contains.
usage: python tools/sass_census.py > profiles/r2_sass_census.txt"""
import collections, os, re, subprocess, sys
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(HERE, "ragmeup_b200", "csrc", "libragmeup_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "ELECT", "BRA.U.ANY", "FFMA2", "FMUL2", "FADD2",
         "HMMA", "LDSM", "MUFU.EX2", "NANOSLEEP"]
print("# SASS instruction census of ragmeup_b200/csrc/libragmeup_b200.so (cuobjdump -sass, sm_100a), one entry per instantiation.")
print("# UTCHMMA = tcgen05.mma (kind::f16 / tf32), UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load,")
print("# SYNCS = mbarrier ops, ELECT = elect.sync (one per issuing role; an ELECT + BRA.U.ANY pair around a tcgen05 / TMA instruction is")
print("# ptxas's single-thread loop for `if (lane == 0)` code -- there must be none), FFMA2/FMUL2/FADD2 = packed fp32 pairs,")
print("# HMMA/LDSM = mma.sync / ldmatrix (the long-sequence attention fallback only).")
print()
blocks = re.split(r"\n\s*Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    if not any(k in name for k in ("scan_rows", "gemm_f16x3", "attention", "select_rescore", "pool_rescore", "exact_scan", "finalize")):
        continue
    ops = collections.Counter()
    n = 0
    for line in blk.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w == "BRA.U.ANY" and op.startswith("BRA.U.ANY")):
                ops[w] += 1
    short = re.sub(r"\(.*", "(...)", name)
    print(short)
    print(f"    {n} instructions: " + ", ".join(f"{k} {v}" for k, v in sorted(ops.items())))
