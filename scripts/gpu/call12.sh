#!/bin/bash
# operator call fused into the step kernel (phase S): parity suite, then A/B against BL_STEP_OP=0
O=gpurun_out/r2c12; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $O/tests_parity.log 2>&1; echo "exit=$?" >> $O/tests_parity.log; tail -4 $O/tests_parity.log
for cfg in "BL_STEP_OP=1" "BL_STEP_OP=0"; do
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single:   $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --lanes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep 1 lane: $(cat $O/q.json)"; tail -2 $O/q.err
done
