"""Row-sharded Lanczos forward + adjoint on a sparse operand (one rank per GPU, torchrun):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/run_row_sharded.py

Checks the sharded result against the single-GPU run of the same problem on rank 0's GPU and
prints one JSON line with the timings."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import parallel, synthetic

rank, world, local_rank = parallel.init_from_env()
n = int(os.environ.get("N", 1_000_000))
K = int(os.environ.get("DEPTH", 40))
dtype = np.float64 if os.environ.get("DTYPE", "f32") == "f64" else np.float32
row, col, data = synthetic.banded_spd_coo(n, bands=5, seed=0)
rng = np.random.default_rng(1)
v = rng.standard_normal(n).astype(dtype)
dalpha, dbeta = rng.standard_normal(K), rng.standard_normal(K - 1)

op = parallel.RowShardedSparseOperator(row, col, n)
alg = bl.lanczos.tridiag(op.callback, K, reortho="full")


def run():
    with parallel.row_sharded():
        ((Qt, (alpha, beta)), _), pull = bl.vjp(alg, op.local_slice(v), data.astype(dtype))
        dv, dp = pull(((None, (dalpha, dbeta)), (None, None)))
    bl.synchronize()
    return alpha, beta, dv, dp


run()
t0 = time.perf_counter()
alpha, beta, dv, dp = run()
t_sharded = time.perf_counter() - t0
res = {"world": world, "n": n, "K": K, "dtype": np.dtype(dtype).name, "sharded_seconds": t_sharded}
if rank == 0:
    full = bl.operators.SparseOperator(row, col, (n, n))
    ref = bl.lanczos.tridiag(full, K, reortho="full")
    ((Qt, (a_ref, b_ref)), _), pull = bl.vjp(ref, v, data.astype(dtype))
    dv_ref, dp_ref = pull(((None, (dalpha, dbeta)), (None, None)))
    bl.synchronize()
    t0 = time.perf_counter()
    ((Qt, (a_ref, b_ref)), _), pull = bl.vjp(ref, v, data.astype(dtype))
    dv_ref, dp_ref = pull(((None, (dalpha, dbeta)), (None, None)))
    bl.synchronize()
    res["single_gpu_seconds"] = time.perf_counter() - t0

    def err(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    chunk = op.chunk
    res.update(err_alpha=err(alpha, a_ref), err_beta=err(beta, b_ref),
               err_dv_local=err(dv.numpy()[: min(chunk, n)], dv_ref.numpy()[:chunk]),
               err_dparams=err(dp, dp_ref.numpy()))
    print(json.dumps(res))
from experiments_lanczos_adjoints_b200 import comm as bl_comm

bl_comm.default().barrier()
bl_comm.shutdown()
