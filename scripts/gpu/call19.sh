#!/bin/bash
O=gpurun_out/r2c19; mkdir -p $O
timeout 300 python scripts/debug_step_op.py 20000 6 f32 > $O/dbg.log 2>&1; echo "exit=$?" >> $O/dbg.log; grep -v Warn $O/dbg.log | grep -A9 "== fused" | grep -v "^  ref\|alpha\|beta"
for d in 4 6 8 12; do
echo "== depth $d"
BL_STEP=2 BL_STEP_DEPTH=$d BL_STEP_L2=9 timeout 300 python scripts/trace_step_kernel.py > $O/trace_$d.json 2>$O/trace.err; head -8 $O/trace_$d.json | grep "phase0"
BL_STEP_DEPTH=$d timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
BL_STEP=2 BL_STEP_DEPTH=$d timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "single: $(cat $O/q.json)"; tail -2 $O/q.err
done
