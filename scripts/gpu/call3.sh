#!/bin/bash
O=gpurun_out/r2c3; mkdir -p $O
timeout 300 python scripts/trace_step_kernel.py 1.9 > $O/trace_coop.json 2> $O/trace_coop.err; echo "trace rc=$?"; cat $O/trace_coop.json
BL_STEP_COOP=0 timeout 300 python scripts/trace_step_kernel.py 1.9 > $O/trace_nocoop.json 2> $O/trace_nocoop.err; echo "trace nocoop rc=$?"; cat $O/trace_nocoop.json
BL_STEP_COOP=0 timeout 300 python bench.py --quick --probes 1 --steps 10 --warmup 3 > $O/bench_nocoop_p1.json 2>/dev/null; echo "nocoop p1 $(cat $O/bench_nocoop_p1.json)"
BL_STEP_COOP=0 BL_STEP_PDL=0 timeout 300 python bench.py --quick --probes 1 --steps 10 --warmup 3 > $O/bench_nocoop_nopdl_p1.json 2>/dev/null; echo "nocoop nopdl p1 $(cat $O/bench_nocoop_nopdl_p1.json)"
timeout 300 python bench.py --quick --probes 1 --steps 1 --warmup 1 > $O/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --quick --probes 1 --steps 1 --warmup 1 > $O/ncu.log 2>&1; echo "ncu rc=$?"
