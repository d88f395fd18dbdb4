#!/bin/bash
O=gpurun_out/r2c63; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "dense or cotangent or golden or arnoldi or batched_initial or pde or wave" 2>&1 | tail -2
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import bench, bench_extra
import experiments_lanczos_adjoints_b200 as bl
row, col, data, _, _ = bench.build_workload()
print(bench_extra.published_recipe(bl, row, col, data, bench.N_ROWS, bench.DEPTH))
PY
