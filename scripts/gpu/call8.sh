#!/bin/bash
# tests with the wide SpMV + K>128, bench, then ncu (launch list + full captures); every ncu command follows a plain run
# of the same command line that exited 0
O=gpurun_out/r2c8; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_all.log 2>&1; echo "exit=$?" >> $O/tests_all.log; tail -3 $O/tests_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; head -c 700 $O/bench.json; echo
BL_SPMV_WIDE=0 timeout 600 python bench.py --quick --steps 5 --warmup 3 > $O/bench_spmv_narrow.json 2>/dev/null; echo "narrow spmv: $(cat $O/bench_spmv_narrow.json)"
timeout 600 python bench.py --quick --steps 5 --warmup 3 > $O/bench_quick.json 2>/dev/null; echo "wide spmv:   $(cat $O/bench_quick.json)"
timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/bench_single.json 2>/dev/null; echo "single: $(cat $O/bench_single.json)"
# --- lockstep lane: launch list, then full captures of the step kernel and the multi-vector SpMV
CMD="python bench.py --quick --steps 1 --warmup 1 --lanes 1 --probes 4"
timeout 300 $CMD > $O/plain_lockstep.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_lockstep.csv $CMD > $O/ncu_ll.log 2>&1; echo "ncu launch list lockstep rc=$?"
timeout 300 $CMD > $O/plain_lockstep2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 294 -c 1 -o $O/prof_step $CMD > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 300 $CMD > $O/plain_lockstep3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_spmv_multi -s 250 -c 2 -o $O/prof_spmv_multi $CMD > $O/ncu_spmv.log 2>&1; echo "ncu spmv multi rc=$?"
# --- one run alone (classic kernels): launch list and full captures of the kernels VERDICT r1 asked for
CMD1="python bench.py --quick --mode streams --probes 1 --steps 1 --warmup 1"
timeout 300 $CMD1 > $O/plain_single.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $O/launches_single.csv $CMD1 > $O/ncu_ls.log 2>&1; echo "ncu launch list single rc=$?"
timeout 300 $CMD1 > $O/plain_single2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_dots_few|k_sell_grad_batch|k_sell_spmv_multi" -s 330 -c 6 -o $O/prof_single_small $CMD1 > $O/ncu_small.log 2>&1; echo "ncu single small rc=$?"
ls -la $O | head -40
