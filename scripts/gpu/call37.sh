#!/bin/bash
# one run alone: operator call + neighbouring-row dots in one launch (k_sell_spmv_dots) against k_dots_few
O=gpurun_out/r2c37; mkdir -p $O
for cfg in "BL_SPMV_DOTS=0" "BL_SPMV_DOTS_THREADS=256" "BL_SPMV_DOTS_THREADS=512" "BL_SPMV_DOTS_THREADS=1024"; do
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single: $(cat $O/q.json)"; tail -2 $O/q.err
done
for cfg in "BL_SPMV_DOTS=0" "BL_SPMV_DOTS_THREADS=512"; do
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 4 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg 4 streams: $(cat $O/q.json)"; tail -2 $O/q.err
  env $cfg timeout 300 python bench.py --quick --mode streams --probes 1 --dtype f64 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg single f64: $(cat $O/q.json)"; tail -2 $O/q.err
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
