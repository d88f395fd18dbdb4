#!/bin/bash
O=gpurun_out/r2c14; mkdir -p $O
timeout 300 python scripts/trace_step_kernel.py > $O/trace_fused.json 2>$O/trace.err; echo "exit=$?"; head -50 $O/trace_fused.json
