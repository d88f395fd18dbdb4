#!/bin/bash
O=gpurun_out/r2c51; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -4 $O/tests.log
timeout 300 python bench.py --quick --steps 6 --warmup 3 2>/dev/null
