#!/bin/bash
# study runs of the scan under gpurun: ring depth / stage size / CTA pairs / ablations (10M x 384)
out=gpurun_out/scan_variants.log
: > $out
run() { echo "== $*" >> $out; env "$@" PROF_CLASSES=1 PROF_ITERS=3 timeout 120 python tools/prof_search.py 2>&1 | grep -v Warning >> $out; }
for q in 64 32 16; do
  run PROF_Q=$q PROF_K=10 RMU_SCAN_KD=2
  run PROF_Q=$q PROF_K=10 RMU_SCAN_KD=1
done
run PROF_Q=64 PROF_K=100 RMU_SCAN_KD=2
run PROF_Q=64 PROF_K=100 RMU_SCAN_KD=1
for ab in 1 2 3; do
  run PROF_Q=64 PROF_K=10 RMU_SCAN_KD=2 RMU_SCAN_ABLATE=$ab RMU_SEARCH_NOFALLBACK=1
done
run PROF_Q=16 PROF_K=10 RMU_SCAN_KD=2 RMU_SCAN_ABLATE=3
run PROF_Q=16 PROF_K=10 RMU_SCAN_KD=1 RMU_SCAN_ABLATE=3
run PROF_Q=16 PROF_K=10 RMU_SCAN_KD=1 RMU_SCAN_ABLATE=3 RMU_SCAN_STAGES=8
run PROF_Q=16 PROF_K=10 RMU_SCAN_KD=1 RMU_SCAN_ABLATE=3 RMU_SCAN_STAGES=10
for kd in 2 1; do
  run PROF_Q=64 PROF_K=10 RMU_SCAN_PAIR=1 RMU_SCAN_KD=$kd
  run PROF_Q=64 PROF_K=10 RMU_SCAN_PAIR=1 RMU_SCAN_KD=$kd RMU_SCAN_ABLATE=3
  run PROF_Q=64 PROF_K=10 RMU_SCAN_PAIR=1 RMU_SCAN_KD=$kd RMU_SCAN_ABLATE=2
done
run PROF_Q=128 PROF_K=10 RMU_SCAN_PAIR=1 RMU_SCAN_KD=2
cat $out
