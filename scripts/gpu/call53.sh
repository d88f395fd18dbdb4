#!/bin/bash
# round-2d profiles (the code as committed at the end of the round): launch list of a lockstep lane, full capture of k_step_tma
# (forward step i = 95).  Every ncu command follows a plain run of the same command line that exited 0.
O=gpurun_out/r2c53; mkdir -p $O
CMD="python bench.py --quick --steps 1 --warmup 1 --lanes 1 --probes 4"
timeout 300 $CMD > $O/plain_lockstep.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_lockstep.csv $CMD > $O/ncu_ll.log 2>&1; echo "ncu launch list lockstep rc=$?"
timeout 300 $CMD > $O/plain_lockstep2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 294 -c 1 -o $O/prof_step $CMD > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
CMD1="python bench.py --quick --mode streams --probes 1 --steps 1 --warmup 1"
timeout 300 $CMD1 > $O/plain_single.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $O/launches_single.csv $CMD1 > $O/ncu_ls.log 2>&1; echo "ncu launch list single rc=$?"
