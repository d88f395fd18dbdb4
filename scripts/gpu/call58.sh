#!/bin/bash
O=gpurun_out/r2c58; mkdir -p $O
timeout 300 python scripts/profile_wave.py 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 300 python scripts/bench_row_sharded_wave.py 2>&1 | tail -2
