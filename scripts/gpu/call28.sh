#!/bin/bash
O=gpurun_out/r2c28; mkdir -p $O
DBG_REPS=3 timeout 400 python scripts/debug_lockstep_diff.py 1000000 20 4 2 f32 2>&1 | grep -v Warn | tee $O/diff.log | head -80
