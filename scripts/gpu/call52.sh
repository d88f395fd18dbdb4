#!/bin/bash
O=gpurun_out/r2c52; mkdir -p $O
for cfg in "BL_STEP_L2=1" "BL_STEP_L2=0" "BL_STEP_L2=3" "BL_STEP_L2=2" "BL_STEP_L2=1 BL_ROWS_L2=0" "BL_STEP_L2=1"; do
  env $cfg timeout 300 python bench.py --quick --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg: $(cut -c1-60 $O/q.json)"
done
