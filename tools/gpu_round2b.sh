#!/bin/bash
# round-2 measurement batch: attention kernels side by side, ncu of the tcgen05 attention and of the scan
mkdir -p gpurun_out
L=gpurun_out/round2b.log
: > $L
echo "== encoder, tcgen05 attention" >> $L
PROF_B=800 PROF_CLASSES=1 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn >> $L
echo "== encoder, mma.sync attention (RMU_ATTN_MODE=3)" >> $L
RMU_ATTN_MODE=3 PROF_B=800 PROF_CLASSES=1 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn >> $L
for cfg in "PROF_Q=64 PROF_K=100" "PROF_Q=64 PROF_K=10" "PROF_Q=16 PROF_K=10" "PROF_Q=128 PROF_K=10" "PROF_Q=64 PROF_K=10 RMU_SCAN_PAIR=1" "PROF_Q=128 PROF_K=10 RMU_SCAN_PAIR=1" "PROF_Q=64 PROF_K=50 PROF_D=768 PROF_N=5000000" "PROF_Q=64 PROF_K=50 PROF_D=768 PROF_N=5000000 RMU_SCAN_PAIR=1"; do
  echo "== scan $cfg" >> $L
  env $cfg PROF_CLASSES=1 PROF_ITERS=3 PROF_METRIC=cosine timeout 200 python tools/prof_search.py 2>&1 | grep -v Warn >> $L
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -o gpurun_out/r2_attention_tc env PROF_B=800 PROF_ITERS=1 python tools/prof_encoder.py > gpurun_out/ncu_att.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_rows -s 1 -c 1 -o gpurun_out/r2_scan_k100 env PROF_Q=64 PROF_K=100 PROF_ITERS=2 PROF_METRIC=cosine python tools/prof_search.py > gpurun_out/ncu_scan.log 2>&1
cat $L
