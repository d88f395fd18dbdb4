#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "slq or hutchinson or estimator or sharded or staged or suitesparse or sparse" 2>&1 | tail -2
PROBES_PER_GPU=64 timeout 300 python scripts/bench_slq.py 2>&1 | grep probes_per_s
python - <<'PY'
import gc, numpy as np, sys
sys.path.insert(0, ".")
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import plan, synthetic
n = 100_000
row, col, data = synthetic.banded_spd_coo(n, 5, seed=0)
probes = (np.random.default_rng(1).integers(0, 2, size=(16, n)) * 2 - 1).astype(np.float32)
for rep in range(3):  # estimators come and go: their plans and pinned staging buffers are given back
    op = bl.operators.SparseOperator(row, col, (n, n))
    est = bl.hutchinson.hutchinson(bl.lanczos.integrand_spd(np.log, 20, op), lambda key: probes)
    v, g = est.value_and_grad(None, data.astype(np.float32))
    del est, op, g
    gc.collect()
    print("rep", rep, float(v), "pinned buffers alive:", len(plan._PINNED))
PY
