#!/bin/bash
timeout 300 python scripts/bench_c4.py 256 2>&1 | grep probes_per_s
PROBES_PER_GPU=128 timeout 300 python scripts/bench_slq.py 2>&1 | grep probes_per_s
timeout 600 python -m pytest tests -m gpu -x -q -k "slq or hutchinson or estimator or sharded" 2>&1 | tail -3
