"""Time of the deferred parameter-cotangent pass (profile class "vjp") of one run and of a lockstep batch at n = 1M, depth 100."""
import sys

import numpy as np

sys.path.insert(0, ".")
import experiments_lanczos_adjoints_b200 as bl  # noqa: E402
from experiments_lanczos_adjoints_b200 import plan as bl_plan, synthetic  # noqa: E402

n, K = 1_000_000, 100
dtype = np.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else np.float64
tight = len(sys.argv) > 2 and sys.argv[2] == "tight"
row, col, data = synthetic.banded_spd_coo(n, 5, seed=0, max_offset=100 if tight else 2000, long_range=0 if tight else 1)
rng = np.random.default_rng(1)
dH1 = synthetic.slq_cotangent_dH(rng.standard_normal(K), rng.standard_normal(K - 1), dtype)
for P in (1, 4):
    op = bl.operators.SparseOperator(row, col, (n, n))
    if P == 1:
        pl = bl_plan.TridiagAdjointPlan(op, K, dtype)
        pl.set_vector(rng.standard_normal(n).astype(dtype))
        pl.set_cotangent(dH1)
    else:
        pl = bl_plan.BatchedTridiagAdjointPlan(op, K, dtype, P)
        pl.set_vectors(rng.standard_normal((P, n)).astype(dtype))
        pl.set_cotangents(np.stack([dH1] * P))
    pl.set_params(data.astype(dtype))
    pl.run()
    pl.stream.synchronize()
    prof = bl_plan.profile(lambda: (pl.run(), pl.stream.synchronize()))
    total = sum(v["ms"] for v in prof.values())
    print(f"P={P} {np.dtype(dtype).name}: vjp class {prof['vjp']['launches']} launches {prof['vjp']['ms']:.3f} ms of {total:.2f} ms", flush=True)
    del pl, op
