#!/bin/bash
# round-2c profiles (final code of the round): launch list of a lockstep lane, full captures of k_step_tma (forward step i = 95,
# operator call inside), k_sell_grad_batch (400 pairs) and k_sell_grad_tma (tight band).  Every ncu command follows a plain run of
# the same command line that exited 0.
O=gpurun_out/r2c36; mkdir -p $O
CMD="python bench.py --quick --steps 1 --warmup 1 --lanes 1 --probes 4"
timeout 300 $CMD > $O/plain_lockstep.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_lockstep.csv $CMD > $O/ncu_ll.log 2>&1; echo "ncu launch list lockstep rc=$?"
timeout 300 $CMD > $O/plain_lockstep2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 294 -c 1 -o $O/prof_step $CMD > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 300 $CMD > $O/plain_lockstep3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_grad_batch -s 1 -c 1 -o $O/prof_grad_batch $CMD > $O/ncu_gb.log 2>&1; echo "ncu grad batch rc=$?"
CMD2="python scripts/time_grad_batch.py f32 tight"
timeout 300 $CMD2 > $O/plain_tight.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_grad_tma -s 3 -c 1 -o $O/prof_grad_tma $CMD2 > $O/ncu_gt.log 2>&1; echo "ncu grad tma rc=$?"
ls -la $O | head -20
