#!/bin/bash
# A/B on one box: libvariant_A.so (reference build) against the current library; the parity tests of the current build run
# first under a short timeout so that a hang costs a minute, not the call
mkdir -p gpurun_out
L=gpurun_out/ab.log
: > $L
D=ragmeup_b200/csrc
cp $D/libragmeup_b200.so $D/libvariant_B.so
timeout 120 python -m pytest tests/test_encoder_gpu.py -x -q 2>&1 | tail -2 >> $L
if ! grep -q "passed" $L || grep -q "failed" $L; then echo "TESTS DID NOT PASS" >> $L; cat $L; exit 1; fi
for rep in 1 2; do
  for v in A B; do
    cp $D/libvariant_$v.so $D/libragmeup_b200.so
    echo "== variant $v (rep $rep)" >> $L
    PROF_B=800 PROF_CLASSES=1 PROF_ITERS=5 timeout 100 python tools/prof_encoder.py 2>&1 | grep -v Warn | tail -1 >> $L
  done
done
cp $D/libvariant_B.so $D/libragmeup_b200.so
cat $L
