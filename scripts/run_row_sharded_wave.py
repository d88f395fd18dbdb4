"""Row-sharded Arnoldi forward + adjoint on the wave-stencil operand (halo exchange), torchrun:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/run_row_sharded_wave.py

Checks against the single-GPU run on rank 0 and prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import parallel

rank, world, local_rank = parallel.init_from_env()
g = int(os.environ.get("GRID", 2048))
K = int(os.environ.get("DEPTH", 10))
dtype = np.float64 if os.environ.get("DTYPE", "f64") == "f64" else np.float32
rng = np.random.default_rng(0)
stencil = bl.operators.WaveStencilOperator.stencil_laplacian(1.0)
xs = np.linspace(0, 1, g)
y0 = np.stack([np.exp(-80 * ((xs[:, None] - 0.4) ** 2 + (xs[None, :] - 0.6) ** 2)),
               0.1 * np.sin(5 * xs)[:, None] * np.ones(g)[None, :]]).astype(dtype)  # fmt: skip
scale = (0.3 + 0.05 * np.sin(6 * xs)[:, None] * np.cos(4 * xs)[None, :]).astype(dtype)
dH = rng.standard_normal((K, K)).astype(dtype)

op = parallel.RowShardedWaveOperator(g, stencil)
alg = bl.arnoldi.hessenberg(op.callback, K, reortho="full")


def run():
    with parallel.row_sharded():
        (Q, H, r, c), pull = bl.vjp(alg, op.local_slice(y0), scale)
        dv, ds = pull((None, dH, None, None))
    bl.synchronize()
    return H.numpy(), dv.numpy(), ds


# operator-level check of the halo exchange: matvec and vjp against the square operand
x_g, lam_g = rng.standard_normal((2, g, g)).astype(dtype), rng.standard_normal((2, g, g)).astype(dtype)
op.callback.bind((scale,), dtype)
op.callback.grad_zero(dtype)
y_loc = op._matvec(bl.asarray(op.local_slice(x_g))).numpy()
z_loc, _ = op._vjp(bl.asarray(op.local_slice(x_g)), bl.asarray(op.local_slice(lam_g)))
z_loc = z_loc.numpy()
(ds_sh,) = op.callback.grad_export(dtype)
sq = bl.operators.WaveStencilOperator(g, stencil)
y_ref = sq(x_g.ravel(), scale).numpy().reshape(2, g, g)
sq.grad_zero(dtype)
z_ref = sq.vjp(bl.asarray(x_g.ravel()), bl.asarray(lam_g.ravel())).numpy().reshape(2, g, g)
(ds_ref,) = sq.grad_export(dtype)
op_err = {
    "matvec": float(np.abs(y_loc - op.local_slice(y_ref)).max() / np.abs(y_ref).max()),
    "vjp_z": float(np.abs(z_loc - op.local_slice(z_ref)).max() / np.abs(z_ref).max()),
    "vjp_dscale": float(np.abs(ds_sh - ds_ref.numpy()).max() / np.abs(ds_ref.numpy()).max()),
}

run()
t0 = time.perf_counter()
H, dv, ds = run()
res = {"world": world, "grid": g, "K": K, "dtype": np.dtype(dtype).name, "sharded_seconds": time.perf_counter() - t0,
       "operator_level_max_err": op_err}
if rank == 0:
    ref = bl.arnoldi.hessenberg(bl.operators.WaveStencilOperator(g, stencil), K, reortho="full")
    (Q0, H0, r0, c0), pull0 = bl.vjp(ref, y0.ravel(), scale)
    dv0, ds0 = pull0((None, dH, None, None))
    bl.synchronize()
    t0 = time.perf_counter()
    (Q0, H0, r0, c0), pull0 = bl.vjp(ref, y0.ravel(), scale)
    dv0, ds0 = pull0((None, dH, None, None))
    bl.synchronize()
    res["single_gpu_seconds"] = time.perf_counter() - t0

    def err(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    dv_ref_local = op.local_slice(dv0.numpy().reshape(2, g, g))
    res.update(err_H=err(H, H0.numpy()), err_dv_local=err(dv, dv_ref_local), err_dscale=err(ds, ds0.numpy()))
    print(json.dumps(res))
from experiments_lanczos_adjoints_b200 import comm as bl_comm

bl_comm.default().barrier()
bl_comm.shutdown()
