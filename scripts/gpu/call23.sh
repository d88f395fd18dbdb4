#!/bin/bash
O=gpurun_out/r2c23; mkdir -p $O
timeout 300 python scripts/debug_step_op.py 20000 6 f32 > $O/dbg.log 2>&1; echo "exit=$?" >> $O/dbg.log; grep -v Warn $O/dbg.log | grep -A12 "== fused" | grep -v "^  ref"
timeout 300 python scripts/debug_step_op.py 30011 5 f64 > $O/dbg64.log 2>&1; echo "exit=$?" >> $O/dbg64.log; grep -v Warn $O/dbg64.log | grep -A12 "== fused" | grep -v "^  ref"
for cfg in "BL_OP_DOTS=1" "BL_OP_DOTS=0"; do
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single:   $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 4 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg 4 streams:   $(cat $O/q.json)"; tail -2 $O/q.err
done
