#!/bin/bash
O=gpurun_out/r2c6; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/tests_all.log 2>&1; echo "exit=$?" >> $O/tests_all.log; tail -4 $O/tests_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -5 $O/bench.err; head -c 6000 $O/bench.json
