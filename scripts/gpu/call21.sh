#!/bin/bash
# whole GPU suite with the operator call inside the step kernel (lockstep batches) and L2 evict_first rows; bench line
O=gpurun_out/r2c21; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --no-extra > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; cat $O/bench.json
