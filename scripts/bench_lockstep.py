"""Lockstep batches on the headline workload (n = 1M, depth 100, fp32): P runs per batch through
`BatchedTridiagAdjointPlan` (multi-vector SpMV + one k_step_tma launch per four runs), device-timed, against P
independent plans; first a parity check of the batch against single runs."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import plan as bl_plan, synthetic

dtype = np.float32 if "--f64" not in sys.argv else np.float64
row, col, data, dalpha, dbeta = bench.build_workload()
n, K = bench.N_ROWS, bench.DEPTH
dH1 = synthetic.slq_cotangent_dH(dalpha, dbeta, dtype)
steps = 5
out = {}
single = bl_plan.TridiagAdjointPlan(bl.operators.SparseOperator(row, col, (n, n)), K, dtype)
single.set_params(data.astype(dtype))
single.set_cotangent(dH1)
for P in [int(a) for a in sys.argv[1:] if a.isdigit()] or [4]:
    op = bl.operators.SparseOperator(row, col, (n, n))
    pl = bl_plan.BatchedTridiagAdjointPlan(op, K, dtype, P)
    V = np.stack([np.random.default_rng(100 + p).standard_normal(n) for p in range(P)]).astype(dtype)
    pl.set_vectors(V)
    pl.set_params(data.astype(dtype))
    pl.set_cotangents(np.stack([dH1] * P))
    for _ in range(3):
        pl.run()
    bl.synchronize()
    e0, e1 = bl.Event(), bl.Event()
    l0 = bl.launch_count()
    e0.record(pl.stream)
    for _ in range(steps):
        pl.run()
    e1.record(pl.stream)
    e1.synchronize()
    ms = e0.elapsed_ms(e1) / steps
    launches = (bl.launch_count() - l0) / steps
    # parity: run 0 and run P-1 of the batch against single runs; the gradient is the sum over the runs
    Hb = pl.H.numpy().reshape(P, K, K)
    dvb = pl.dv.numpy()
    gb = pl.grads[0].numpy()
    gsum = np.zeros_like(gb, dtype=np.float64)
    errs = {"H": 0.0, "dv": 0.0}
    for p in range(P):
        single.set_vector(V[p])
        single.run()
        Hs, dvs = single.H.numpy(), single.dv.numpy()
        gsum += single.grads[0].numpy()
        errs["H"] = max(errs["H"], float(np.abs(Hb[p] - Hs).max() / np.abs(Hs).max()))
        errs["dv"] = max(errs["dv"], float(np.linalg.norm(dvb[p] - dvs) / np.linalg.norm(dvs)))
    errs["grad_sum"] = float(np.linalg.norm(gb - gsum) / np.linalg.norm(gsum))
    out[P] = {"ms_per_batch": ms, "ms_per_run": ms / P, "krylov_steps_per_s": P * K / (ms * 1e-3),
              "launches_per_batch": launches, "vs_single_runs": errs}
    print(P, json.dumps(out[P]), flush=True)
    del pl, op
print(json.dumps(out))
