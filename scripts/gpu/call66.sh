#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "gram or gp or lockstep_probe or logml or marginal or cholesky" 2>&1 | tail -2
timeout 300 python scripts/bench_gp_logml.py 2>&1 | tail -1 | cut -c1-420
