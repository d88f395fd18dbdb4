#!/bin/bash
O=gpurun_out/r2c25; mkdir -p $O
echo "== fused, 2 lanes, 24 reps, K=12"; DBG_REPS=24 timeout 300 python scripts/debug_lockstep_op.py child /tmp/x.npz 1000000 12 4 2 f32 2>&1 | grep -v Warn | tail -4
echo "== fused, 2 lanes, 8 reps, K=30"; DBG_REPS=8 timeout 300 python scripts/debug_lockstep_op.py child /tmp/x.npz 1000000 30 4 2 f32 2>&1 | grep -v Warn | tail -4
echo "== separate, 3 lanes, K=30"; BL_STEP_OP=0 DBG_REPS=6 timeout 300 python scripts/debug_lockstep_op.py child /tmp/x.npz 1000000 30 4 3 f32 2>&1 | grep -v Warn | tail -4
echo "== fused, 3 lanes, K=30"; DBG_REPS=6 timeout 300 python scripts/debug_lockstep_op.py child /tmp/x.npz 1000000 30 4 3 f32 2>&1 | grep -v Warn | tail -4
