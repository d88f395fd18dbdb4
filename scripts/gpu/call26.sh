#!/bin/bash
O=gpurun_out/r2c26; mkdir -p $O
BL_STEP_OP=1 DBG_REPS=2 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/debug_lockstep_op.py child /tmp/x.npz 65536 12 4 2 f32 > $O/memcheck.log 2>&1; echo "exit=$?"; grep -v "^$" $O/memcheck.log | tail -40
