#!/bin/bash
# full capture (with source) of one fused step kernel launch of a single run
O=gpurun_out/r2c16; mkdir -p $O
CMD1="python bench.py --quick --mode streams --probes 1 --steps 1 --warmup 1"
timeout 300 $CMD1 > $O/plain_single.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 260 -c 1 -o $O/prof_step_fused $CMD1 > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
ls -la $O
