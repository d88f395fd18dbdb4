#!/bin/bash
O=gpurun_out/r2c18; mkdir -p $O
for h in 9 25 41 57 13; do
echo "== BL_STEP_L2=$h (8: S-only stamp, 16: no gathers, 32: no value loads, 4: no L2 prefetch of values)"
BL_STEP=2 BL_STEP_L2=$h timeout 300 python scripts/trace_step_kernel.py > $O/trace_$h.json 2>$O/trace.err; head -8 $O/trace_$h.json | grep "phase0"
done
for h in 9 25 41; do
echo "== lockstep batch of 4, BL_STEP_L2=$h"
TRACE_PROBES=4 BL_STEP_L2=$h timeout 300 python scripts/trace_step_kernel.py > $O/trace4_$h.json 2>$O/trace.err; head -12 $O/trace4_$h.json | grep "phase"
done
