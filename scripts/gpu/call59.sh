#!/bin/bash
# final code: throughput by runs per batch x lanes (bench.py --quick), one JSON line each
O=gpurun_out/r2c59; mkdir -p $O; : > $O/lanes.jsonl
for cfg in "--probes 4 --lanes 1" "--probes 2 --lanes 2" "--probes 4 --lanes 2" "--probes 4 --lanes 3" "--probes 3 --lanes 2" "--mode streams --probes 1" "--mode streams --probes 4"; do
  timeout 300 python bench.py --quick --steps 6 --warmup 3 $cfg > $O/q.json 2>$O/q.err && echo "{\"args\": \"$cfg\", \"result\": $(cat $O/q.json)}" >> $O/lanes.jsonl
done
BL_BENCH_STAGGER=0 timeout 300 python bench.py --quick --steps 6 --warmup 3 > $O/q.json 2>$O/q.err && echo "{\"args\": \"--probes 4 --lanes 2, BL_BENCH_STAGGER=0\", \"result\": $(cat $O/q.json)}" >> $O/lanes.jsonl
BL_STEP_OP=0 timeout 300 python bench.py --quick --steps 6 --warmup 3 > $O/q.json 2>$O/q.err && echo "{\"args\": \"--probes 4 --lanes 2, BL_STEP_OP=0\", \"result\": $(cat $O/q.json)}" >> $O/lanes.jsonl
cat $O/lanes.jsonl | cut -c1-200
