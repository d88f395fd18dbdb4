#!/bin/bash
O=gpurun_out/r2c42; mkdir -p $O
timeout 300 python scripts/bench_c4.py 1024 2>&1 | grep probes_per_s
timeout 600 python bench.py --dtype f64 --no-extra > $O/bench_f64.json 2> $O/bench_f64.err; echo "f64 exit=$?"; python -c "
import json; d=json.load(open('$O/bench_f64.json')); print('f64 value', d['value'], 'e2e', d['e2e']['value'], 'single', d['config']['single_probe'], 'frac', d['roofline']['frac'], 'whole', d['roofline']['whole_step']['frac'])"
timeout 900 python -m pytest tests -m gpu -x -q -k "sparse or clone or concurrent or lanes or slq" 2>&1 | tail -2
