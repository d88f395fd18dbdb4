#!/bin/bash
O=gpurun_out/r2c31; mkdir -p $O
timeout 600 python scripts/check_grad_tma.py 2>&1 | grep -v Warn | tail -6
for cfg in "BL_GRAD_TMA=0" "BL_GRAD_TMA=1"; do
  env $cfg timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single: $(cat $O/q.json)"; tail -2 $O/q.err
done
timeout 900 python -m pytest tests -m gpu -x -q -k "sparse or parity or suite or slq or lockstep" > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
