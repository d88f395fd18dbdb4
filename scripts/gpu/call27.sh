#!/bin/bash
# bisect the 2-lane failure of the operator call inside the step kernel (n = 1M, K = 30, 4 runs per batch)
O=gpurun_out/r2c27; mkdir -p $O
run() { echo "== $1"; shift; env "$@" DBG_REPS=5 timeout 200 python scripts/debug_lockstep_op.py child /tmp/x.npz 1000000 30 4 ${LANES:-2} f32 2>&1 | grep -v Warn | tail -${TAILN:-3}; }
run "F: fused 2 lanes (partials0 fix only)" A=1
run "A: fused 2 lanes, BL_STEP_PDL=0" BL_STEP_PDL=0
run "T: fused 2 lanes, no early trigger" BL_STEP_L2=129
LANES=1 run "B: fused 1 lane, 1 block/SM" DBG_BPS=1
run "C: fused 2 lanes, 2 blocks/SM" DBG_BPS=2
run "D: fused forward only" BL_STEP_OP_SIDES=1
run "E: fused adjoint only" BL_STEP_OP_SIDES=2
run "G: fused 2 lanes, gathers via ld.global.nc" BL_STEP_L2=65
