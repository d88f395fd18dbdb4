#!/bin/bash
O=gpurun_out/r2c4; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gram_tensor.py -x -q -k "slq_estimator or lockstep or sparse_tridiag or golden" > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
if grep -q "exit=0" $O/tests.log; then
  timeout 600 python scripts/bench_lockstep.py 1 2 4 8 > $O/lockstep.log 2>&1; echo "lockstep rc=$?"; cat $O/lockstep.log | head -8
  BL_STEP=2 timeout 300 python bench.py --quick --probes 1 --steps 10 --warmup 3 > $O/bench_step2_p1.json 2>/dev/null; echo "step=2 p1 $(cat $O/bench_step2_p1.json)"
  BL_STEP=0 timeout 600 python scripts/bench_lockstep.py 4 > $O/lockstep_nostep.log 2>&1; echo "lockstep BL_STEP=0:"; head -1 $O/lockstep_nostep.log
  timeout 600 python scripts/bench_lockstep.py --f64 4 > $O/lockstep_f64.log 2>&1; echo "lockstep f64:"; head -1 $O/lockstep_f64.log
fi
