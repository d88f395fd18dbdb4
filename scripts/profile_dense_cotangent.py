"""Per-kernel-class times of the published recipe's adjoint (dense N(0,1) cotangent on every output, benchmark.py:95-122)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import experiments_lanczos_adjoints_b200 as bl  # noqa: E402
from experiments_lanczos_adjoints_b200 import plan as bl_plan  # noqa: E402

row, col, data, _, _ = bench.build_workload()
n, K, dtype = bench.N_ROWS, bench.DEPTH, np.float32
rng = np.random.default_rng(K)
v, p = bl.asarray(rng.standard_normal(n).astype(dtype)), bl.asarray(data.astype(dtype))
cot = ((bl.asarray(rng.standard_normal((K, n)).astype(dtype)), (rng.standard_normal(K), rng.standard_normal(K - 1))),
       (bl.asarray(rng.standard_normal(n).astype(dtype)), float(rng.standard_normal())))
for name, assume in (("general", False), ("symmetric", None)):
    op = bl.operators.SparseOperator(row, col, (n, n))
    alg = bl.lanczos.tridiag(op, K, reortho="full", assume_symmetric=assume)
    _, pull = bl.vjp(alg, v, p)
    pull(cot)
    bl.synchronize()
    prof = bl_plan.profile(lambda: (pull(cot), bl.synchronize()))
    print(name, json.dumps({k: (c["launches"], round(c["ms"], 3), round(c["algorithmic_bytes"] / max(c["ms"], 1e-9) / 1e6)) for k, c in prof.items()}),
          "sum", round(sum(c["ms"] for c in prof.values()), 2))
