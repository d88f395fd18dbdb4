#!/bin/bash
O=gpurun_out/r2c17; mkdir -p $O
for cfg in "BL_STEP_L2=0" "BL_STEP_L2=1" "BL_STEP_L2=2" "BL_STEP_L2=3"; do
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single:   $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
done
BL_STEP_L2=3 timeout 300 python scripts/trace_step_kernel.py > $O/trace_l2.json 2>$O/trace.err; head -12 $O/trace_l2.json
