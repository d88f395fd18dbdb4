#!/bin/bash
# C4 through the estimator API (hutchinson_sharded -> probe_lockstep_sum), lanes in phase against half a cycle apart
for cfg in "BL_PROBE_STAGGER=0" "BL_PROBE_STAGGER=1"; do
  env $cfg PROBES_PER_GPU=128 timeout 300 python scripts/bench_slq.py 2>&1 | grep probes_per_s | sed "s/^/$cfg: /"
done
timeout 600 python -m pytest tests -m gpu -x -q -k "slq or hutchinson or estimator or spmv or operator_call" 2>&1 | tail -3
