#!/bin/bash
O=gpurun_out/r2c20; mkdir -p $O
for cfg in "BL_ROWS_L2=0" "BL_ROWS_L2=1"; do
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single classic: $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 4 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg 4 streams classic: $(cat $O/q.json)"; tail -2 $O/q.err
done
for cfg in "BL_BENCH_LANES=2" "BL_BENCH_LANES=3"; do
  env $cfg timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
done
timeout 300 python bench.py --quick --probes 2 --lanes 4 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "2x4 lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
timeout 300 python bench.py --quick --probes 3 --lanes 3 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "3x3 lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
timeout 300 python bench.py --quick --dtype f64 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "f64 lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
BL_STEP_OP=0 timeout 300 python bench.py --quick --dtype f64 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "f64 lockstep unfused: $(cat $O/q.json)"; tail -2 $O/q.err
