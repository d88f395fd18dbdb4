#!/bin/bash
export BL_NVCC_EXTRA=-DBL_STEP_CYCLES
for cfg in "BL_STEP_L2=1" "BL_STEP_L2=49"; do echo "== $cfg"; env $cfg timeout 300 python scripts/debug_phase_s_cycles.py 2>&1 | grep "forward"; done
