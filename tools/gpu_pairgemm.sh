#!/bin/bash
# CTA-pair GEMM against the single-CTA kernel: encoder parity tests, per-class times, headline bench
mkdir -p gpurun_out
L=gpurun_out/pairgemm.log
: > $L
echo "== encoder tests (pair default)" >> $L
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q 2>&1 | tail -4 >> $L
echo "== classes, RMU_GEMM_PAIR=0" >> $L
RMU_GEMM_PAIR=0 PROF_B=800 PROF_CLASSES=1 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn >> $L
echo "== classes, pair (default)" >> $L
PROF_B=800 PROF_CLASSES=1 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn >> $L
echo "== classes, pair BN=192" >> $L
RMU_GEMM_BN=192 PROF_B=800 PROF_CLASSES=1 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn >> $L
echo "== bench" >> $L
timeout 600 python bench.py > gpurun_out/bench_pairgemm.json 2>> $L
python tools/show_bench.py gpurun_out/bench_pairgemm.json >> $L 2>&1
cat $L
