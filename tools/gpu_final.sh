#!/bin/bash
# validation batch: the whole GPU suite, smoke(), every bench configuration, the reference arm, launch list of the headline
mkdir -p gpurun_out
L=gpurun_out/final.log
: > $L
echo "== pytest -m gpu" >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $L
for c in headline c2 c3 c4 c5; do
  echo "== bench $c" >> $L
  timeout 900 python bench.py --config $c > gpurun_out/bench_$c.json 2>> $L
  python tools/show_bench.py gpurun_out/bench_$c.json >> $L 2>&1
done
echo "== reference arm" >> $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> $L
tail -c 600 gpurun_out/bench_reference.json >> $L
echo "== launch list (ncu, headline, 2 steps)" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300 >> $L
cat $L
