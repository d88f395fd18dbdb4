"""L lockstep batches of P runs in flight on separate streams (headline workload): does one batch's kernels fill the
other's reductions?  usage: bench_lockstep_lanes.py P L blocks_per_sm"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import device as bl_dev, plan as bl_plan, synthetic

P, L, bps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dtype = np.float32
row, col, data, dalpha, dbeta = bench.build_workload()
n, K = bench.N_ROWS, bench.DEPTH
dH1 = synthetic.slq_cotangent_dH(dalpha, dbeta, dtype)
bl.set_blocks_per_sm(bps)
plans = []
for l in range(L):
    op = bl.operators.SparseOperator(row, col, (n, n))
    pl = bl_plan.BatchedTridiagAdjointPlan(op, K, dtype, P, stream=bl_dev.Stream())
    pl.set_vectors(np.stack([np.random.default_rng(100 + l * P + p).standard_normal(n) for p in range(P)]).astype(dtype))
    pl.set_params(data.astype(dtype))
    pl.set_cotangents(np.stack([dH1] * P))
    plans.append(pl)
steps = 5
for _ in range(2):
    for pl in plans:
        pl.run()
bl.synchronize()
e0, ends = bl.Event(), [bl.Event() for _ in plans]
e0.record(plans[0].stream)
for _ in range(steps):
    for pl in plans:
        pl.run()
for pl, e in zip(plans, ends):
    e.record(pl.stream)
for e in ends:
    e.synchronize()
ms = max(e0.elapsed_ms(e) for e in ends) / steps
print(json.dumps({"P": P, "lanes": L, "blocks_per_sm": bps, "ms_per_step": ms, "ms_per_run": ms / (P * L),
                  "krylov_steps_per_s": P * L * K / (ms * 1e-3)}))
