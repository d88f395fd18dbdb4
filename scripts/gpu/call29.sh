#!/bin/bash
export BL_NVCC_EXTRA=-DBL_STEP_DEBUG
O=gpurun_out/r2c29; mkdir -p $O
DBG_REPS=4 timeout 400 python scripts/debug_lockstep_diff.py 1000000 20 4 2 f32 2>&1 | grep -v Warn | tee $O/diff.log | grep -v "^  \|H run" | cut -c1-400 | head -80
