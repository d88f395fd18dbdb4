"""The headline operand ("one large operator": sparse SPD, n = 1M, ~11 nnz/row, Lanczos depth 100, forward +
adjoint) with its ROWS sharded over the ranks on the native peer-memory route: all-gather of the Lanczos
vector and the dot-product reductions are single kernels over NVLink peer memory (no NCCL, no host callback).
Strong scaling: the total work is fixed.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_row_sharded_sparse.py

Inputs are resident on the devices; the timed region is REPS forward+adjoint sweeps between barriers, max over
ranks.  Rank 0 also times the unsharded operand on its own GPU and checks coefficients and cotangents."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import parallel, synthetic

rank, world, local_rank = parallel.init_from_env()
n = int(os.environ.get("N", 1_000_000))
K = int(os.environ.get("DEPTH", 100))
reps = int(os.environ.get("REPS", 5))
dtype = np.float64 if os.environ.get("DTYPE", "f32") == "f64" else np.float32
row, col, data = synthetic.banded_spd_coo(n, bands=5, seed=0)
rng = np.random.default_rng(1)
v = rng.standard_normal(n).astype(dtype)
dalpha, dbeta = rng.standard_normal(K).astype(dtype), rng.standard_normal(K - 1).astype(dtype)

from experiments_lanczos_adjoints_b200 import comm as bl_comm

group = bl_comm.default()  # the library's socket communicator (parallel.init_from_env); collectives are NCCL in the library


def barrier():
    bl.synchronize()
    group.barrier()


comm = parallel.PeerComm()
op = parallel.RowShardedSparseOperator(row, col, n, comm=comm)
alg = bl.lanczos.tridiag(op.callback, K, reortho="full")
v_loc = bl.asarray(op.local_slice(v))
params = tuple(bl.asarray(p.astype(dtype)) for p in op.local_params(data))


def sweep():
    with parallel.row_sharded(comm=comm):
        ((_, (alpha, beta)), _), pull = bl.vjp(alg, v_loc, *params)
        dv, dpa, _ = pull(((None, (dalpha, dbeta)), (None, None)))
    return alpha, beta, dv, dpa


for _ in range(2):
    alpha, beta, dv, dpa = sweep()
barrier()
e0, e1 = bl.Event(), bl.Event()
e0.record()
for _ in range(reps):
    alpha, beta, dv, dpa = sweep()
e1.record()
e1.synchronize()
ms = e0.elapsed_ms(e1) / reps
barrier()
ms = float(group.allreduce_host(np.array(ms), op="max"))
res = {"config": "sparse SPD operand, rows sharded (peer-memory route), Lanczos fwd+adjoint", "n": n, "nnz": len(data),
       "K": K, "dtype": np.dtype(dtype).name, "world": world, "sharded_ms": ms, "krylov_steps_per_s": K / (ms * 1e-3),
       "timed_out": bool(comm.timed_out())}
grad = op.assemble_grad(dpa)
if rank == 0 and os.environ.get("SINGLE", "1") == "1":
    full = bl.lanczos.tridiag(bl.operators.SparseOperator(row, col, (n, n)), K, reortho="full")
    v0, p0 = bl.asarray(v), bl.asarray(data.astype(dtype))

    def single():
        ((_, (a, b)), _), pull = bl.vjp(full, v0, p0)
        return a, b, pull(((None, (dalpha, dbeta)), (None, None)))

    for _ in range(2):
        a0, b0, (dv0, dp0) = single()
    bl.synchronize()
    e0.record()
    for _ in range(reps):
        a0, b0, (dv0, dp0) = single()
    e1.record()
    e1.synchronize()
    res["single_gpu_ms"] = e0.elapsed_ms(e1) / reps
    res["speedup"] = res["single_gpu_ms"] / ms

    def err(x, y):
        x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
        return float(np.linalg.norm(x - y) / np.linalg.norm(y))

    lo = rank * op.chunk
    hi = min(n, lo + op.chunk)
    res.update(err_alpha=err(alpha, a0), err_beta=err(beta, b0), err_dv_local=err(dv.numpy()[: hi - lo], dv0.numpy()[lo:hi]),
               err_dparams=err(grad, dp0.numpy()))
if rank == 0:
    print(json.dumps(res))
group.barrier()
bl_comm.shutdown()
