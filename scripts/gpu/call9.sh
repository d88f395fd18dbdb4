#!/bin/bash
O=gpurun_out/r2c9; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "sparse_tridiag or golden or slq_estimator" > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -2 $O/tests.log
for cfg in "BL_SPMV_W=0" "BL_SPMV_W=12" "BL_SPMV_WIDE=0"; do
  env $cfg timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>/dev/null; echo "$cfg lockstep: $(cat $O/q.json)"
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>/dev/null; echo "$cfg single:   $(cat $O/q.json)"
done
