#!/bin/bash
O=gpurun_out/r2c13; mkdir -p $O
timeout 300 python scripts/debug_step_op.py 20000 4 f32 > $O/dbg.log 2>&1; echo "exit=$?" >> $O/dbg.log; grep -v Warn $O/dbg.log | grep -A8 "== fused"
BL_STEP_PDL=0 timeout 300 python scripts/debug_step_op.py 20000 4 f32 > $O/dbg_nopdl.log 2>&1; echo "exit=$?" >> $O/dbg_nopdl.log; grep -v Warn $O/dbg_nopdl.log | grep -A8 "== fused"
