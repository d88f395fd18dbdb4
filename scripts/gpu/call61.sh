#!/bin/bash
# one run alone (final code): full captures of k_xdots_tma and k_combine_tma at forward step i = 95
O=gpurun_out/r2c61; mkdir -p $O
CMD1="python bench.py --quick --mode streams --probes 1 --steps 1 --warmup 1"
timeout 300 $CMD1 > $O/plain_single.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_xdots_tma -s 294 -c 1 -o $O/prof_xdots $CMD1 > $O/ncu_x.log 2>&1; echo "ncu xdots rc=$?"
timeout 300 $CMD1 > $O/plain_single2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_combine_tma -s 294 -c 1 -o $O/prof_combine $CMD1 > $O/ncu_c.log 2>&1; echo "ncu combine rc=$?"
