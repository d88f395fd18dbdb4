"""Where the SLQ log-determinant + gradient of the GP path spends its time (lockstep batch, n = 36 560)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import plan as bl_plan

n, d, K, P = int(os.environ.get("GP_N", 36560)), 9, 10, int(os.environ.get("PROBES", 10))
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d))
probes = (rng.integers(0, 2, size=(P, n)) * 2 - 1).astype(np.float32)
op = bl.operators.GramOperator(X, kind="matern32")
est = bl.hutchinson.hutchinson(bl.lanczos.integrand_spd(np.log, K, op), lambda key: probes)
params = (np.full(d, 1.0, np.float32), np.float32(0.5), np.float32(0.3))
est.value_and_grad(None, *params)
bl.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    est.value_and_grad(None, *params)
bl.synchronize()
wall = (time.perf_counter() - t0) / 3
prof = bl_plan.profile(lambda: est.value_and_grad(None, *params))
print(json.dumps({"wall_ms": wall * 1e3, "classes": {k: (c["launches"], round(c["ms"], 3)) for k, c in prof.items()}}))
