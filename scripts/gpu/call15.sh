#!/bin/bash
O=gpurun_out/r2c15; mkdir -p $O
timeout 300 python scripts/debug_step_op.py 20000 6 f32 > $O/dbg.log 2>&1; echo "exit=$?" >> $O/dbg.log; grep -v Warn $O/dbg.log | grep -A12 "== fused" | grep -v "^  ref\|alpha\|beta"
timeout 300 python scripts/debug_step_op.py 30011 5 f64 > $O/dbg64.log 2>&1; echo "exit=$?" >> $O/dbg64.log; grep -v Warn $O/dbg64.log | grep -A12 "== fused" | grep -v "^  ref\|alpha\|beta"
timeout 300 python scripts/trace_step_kernel.py > $O/trace_fused.json 2>$O/trace.err; echo "exit=$?"; head -12 $O/trace_fused.json; grep "body_us_mean\|entry_to_entry_us_mean" $O/trace_fused.json
for cfg in "BL_STEP_OP=1" ; do
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single:   $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --lanes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep 1 lane: $(cat $O/q.json)"; tail -2 $O/q.err
done
