"""Debug build (-DBL_STEP_DEBUG): where block 0 / warp 0 spends its cycles inside phase S of a lockstep batch of four."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import experiments_lanczos_adjoints_b200 as bl  # noqa: E402
from experiments_lanczos_adjoints_b200 import _lib, plan as bl_plan, synthetic  # noqa: E402

row, col, data, dalpha, dbeta = bench.build_workload()
n, K, P, dtype = bench.N_ROWS, 20, 4, np.float32
pl = bl_plan.BatchedTridiagAdjointPlan(bl.operators.SparseOperator(row, col, (n, n)), K, dtype, P)
pl.set_vectors(np.random.default_rng(0).standard_normal((P, n)).astype(dtype))
pl.set_params(data.astype(dtype))
pl.set_cotangents(np.stack([synthetic.slq_cotangent_dH(dalpha[:K], dbeta[: K - 1], dtype)] * P))
pl.run()
bl.synchronize()
lib = _lib.load()
out = (C.c_ulonglong * 16)()
lib.bl_step_debug_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
lib.bl_step_debug_read(out, 1)
pl.forward()
bl.synchronize()
lib.bl_step_debug_read(out, 1)
names = ["group set-up", "ring release + wait", "loads + FMAs", "stores", "tail", "-"]
launches = K
print("forward, per launch (block 0, warp 0), us at 1.9 GHz:", {nm: round(out[9 + k] / launches / 1.9e3, 2) for k, nm in enumerate(names)})
