#!/bin/bash
O=gpurun_out/r2c60; mkdir -p $O
run() { echo -n "$1: "; shift; env "$@" > $O/q.json 2>$O/q.err; cut -c1-110 $O/q.json; tail -1 $O/q.err | cut -c1-200; }
run "f64 lockstep" X=1 timeout 300 python bench.py --quick --dtype f64 --steps 4 --warmup 3
run "f64 single" X=1 timeout 300 python bench.py --quick --dtype f64 --mode streams --probes 1 --steps 4 --warmup 3
run "BL_STEP=2 single" BL_STEP=2 timeout 300 python bench.py --quick --mode streams --probes 1 --steps 4 --warmup 3
run "BL_OP_DOTS=1 single" BL_OP_DOTS=1 timeout 300 python bench.py --quick --mode streams --probes 1 --steps 4 --warmup 3
run "BL_SPMV_DOTS=1 single" BL_SPMV_DOTS=1 timeout 300 python bench.py --quick --mode streams --probes 1 --steps 4 --warmup 3
run "BL_STEP=0 lockstep" BL_STEP=0 timeout 300 python bench.py --quick --steps 4 --warmup 3
run "general loops single" BL_SYMMETRIC_FORWARD=0 BL_SYMMETRIC_ADJOINT=0 timeout 300 python bench.py --quick --mode streams --probes 1 --steps 3 --warmup 2
BL_OP_DOTS=1 timeout 600 python -m pytest tests -m gpu -x -q -k "headline or sparse_tridiag or golden" 2>&1 | tail -1
BL_STEP=2 timeout 600 python -m pytest tests -m gpu -x -q -k "headline or sparse_tridiag or golden" 2>&1 | tail -1
