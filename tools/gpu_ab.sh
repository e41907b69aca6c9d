#!/bin/bash
# A/B on one box: libvariant_A.so (reference build) against the current library, interleaved twice
mkdir -p gpurun_out
L=gpurun_out/ab.log
: > $L
D=ragmeup_b200/csrc
cp $D/libragmeup_b200.so $D/libvariant_B.so
for rep in 1 2; do
  for v in A B; do
    cp $D/libvariant_$v.so $D/libragmeup_b200.so
    echo "== variant $v (rep $rep)" >> $L
    PROF_B=800 PROF_CLASSES=1 PROF_ITERS=5 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn | tail -1 >> $L
  done
done
cp $D/libvariant_B.so $D/libragmeup_b200.so
timeout 300 python -m pytest tests/test_encoder_gpu.py -x -q 2>&1 | tail -2 >> $L

cat $L
