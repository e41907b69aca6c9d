#!/bin/bash
# batch after the elect.sync change: microbenchmark, parity tests, encoder / scan profiles, headline bench
mkdir -p gpurun_out
L=gpurun_out/elect.log
: > $L
(cd tools/micro && ./mma_cost) >> $L 2>&1
echo "== encoder classes" >> $L
PROF_B=800 PROF_CLASSES=1 timeout 200 python tools/prof_encoder.py 2>&1 | grep -v Warn >> $L
echo "== scan Q=64 k=10" >> $L
PROF_Q=64 PROF_K=10 PROF_CLASSES=1 PROF_ITERS=3 PROF_METRIC=cosine timeout 200 python tools/prof_search.py 2>&1 | grep -v Warn >> $L
echo "== tests" >> $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
echo "== bench" >> $L
timeout 600 python bench.py > gpurun_out/bench_elect.json 2>> $L
python tools/show_bench.py gpurun_out/bench_elect.json >> $L 2>&1
cat $L
