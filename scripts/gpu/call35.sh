#!/bin/bash
O=gpurun_out/r2c35; mkdir -p $O
for cfg in "BL_BENCH_STAGGER=0" "BL_BENCH_STAGGER=1" "BL_BENCH_STAGGER=0" "BL_BENCH_STAGGER=1"; do
  env $cfg timeout 300 python bench.py --quick --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
done
BL_BENCH_STAGGER=1 timeout 300 python bench.py --quick --lanes 3 --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "stagger 3 lanes: $(cat $O/q.json)"; tail -2 $O/q.err
BL_BENCH_STAGGER=1 timeout 300 python bench.py --quick --dtype f64 --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "stagger f64: $(cat $O/q.json)"; tail -2 $O/q.err
BL_BENCH_STAGGER=0 timeout 300 python bench.py --quick --dtype f64 --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "in phase f64: $(cat $O/q.json)"; tail -2 $O/q.err
