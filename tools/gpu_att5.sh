#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/att5.log
: > $L
echo "== encoder tests" >> $L
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q 2>&1 | tail -4 >> $L
echo "== classes" >> $L
PROF_B=800 PROF_CLASSES=1 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn >> $L
echo "== ragged" >> $L
timeout 200 python tools/att_ragged.py 2>&1 | tail -6 >> $L
RMU_ATTN_TRACE=1 PROF_B=800 timeout 200 python tools/prof_encoder.py > gpurun_out/prof_trace.log 2>&1
cat $L
