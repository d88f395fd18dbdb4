#!/bin/bash
O=gpurun_out/r2c5; mkdir -p $O
for cfg in "4 2 1" "4 2 2" "2 2 1" "4 1 1" "2 4 1" "4 3 1"; do
  timeout 300 python scripts/bench_lockstep_lanes.py $cfg >> $O/lanes.log 2>$O/lanes.err; tail -1 $O/lanes.log
done
