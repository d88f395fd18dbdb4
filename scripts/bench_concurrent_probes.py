"""Several independent Lanczos forward + adjoint runs (Hutchinson probes) on ONE GPU, each on its own stream
with its own plan: while one run sits in a kernel's ramp / grid-wide reduction tail, the other runs' kernels use
the memory system.  Prints one JSON line per probe count: Krylov steps/s over all probes.

    python scripts/bench_concurrent_probes.py [max_probes]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import device as dev
from experiments_lanczos_adjoints_b200 import plan as bl_plan
from experiments_lanczos_adjoints_b200 import synthetic

n, K, dtype = int(os.environ.get("N", 1_000_000)), int(os.environ.get("DEPTH", 100)), np.float32
row, col, data = synthetic.banded_spd_coo(n, bands=5, seed=0)
rng = np.random.default_rng(0)
dalpha, dbeta = rng.standard_normal(K), rng.standard_normal(K - 1)
dH = synthetic.slq_cotangent_dH(dalpha, dbeta, dtype)
max_p = int(sys.argv[1]) if len(sys.argv) > 1 else 3
plans = []
for p in range(max_p):
    op = bl.operators.SparseOperator(row, col, (n, n))
    pl = bl_plan.TridiagAdjointPlan(op, K, dtype, stream=dev.Stream())
    pl.set_vector(np.random.default_rng(100 + p).standard_normal(n).astype(dtype))
    pl.set_params(data.astype(dtype))
    pl.set_cotangent(dH)
    plans.append(pl)
bl.synchronize()
ref = None
for P in range(1, max_p + 1):
    for _ in range(3):
        for pl in plans[:P]:
            pl.run()
    bl.synchronize()
    steps = 5
    t0 = time.perf_counter()
    for _ in range(steps):
        for pl in plans[:P]:
            pl.run()
    bl.synchronize()
    dt = (time.perf_counter() - t0) / steps
    a, b = plans[0].coefficients()
    g = plans[0].grads[0].numpy(plans[0].stream)
    if ref is None:
        ref = (a, b, g)
    same = bool(np.array_equal(a, ref[0]) and np.array_equal(b, ref[1]) and np.array_equal(g, ref[2]))
    print(json.dumps({"probes_in_flight": P, "ms_per_round": dt * 1e3, "krylov_steps_per_s": P * K / dt,
                      "probe0_bitwise_equal_to_solo_run": same}), flush=True)
