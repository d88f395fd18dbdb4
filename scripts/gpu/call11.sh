#!/bin/bash
# ceilings of the streaming pattern: stand-alone ring probe (stage granularity, L2 prefetch ahead of the ring)
O=gpurun_out/r2c11; mkdir -p $O
timeout 300 scripts/microbench_hbm > $O/hbm2.log 2>&1; echo "exit=$?" >> $O/hbm2.log; cat $O/hbm2.log
