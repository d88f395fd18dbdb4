#!/bin/bash
O=gpurun_out/r2c64; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
