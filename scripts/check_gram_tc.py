"""Tensor-core Gram path vs the ALU path and a float64 NumPy evaluation (GPU check + timings)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from oracle import operators as oops

out = {}
rng = np.random.default_rng(0)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    bl.synchronize()
    e0, e1 = bl.Event(), bl.Event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / reps


for n, d, kind in [(300, 9, "matern32"), (1000, 9, "rbf"), (777, 3, "matern12"), (2500, 16, "matern32"), (513, 20, "rbf")]:
    X = rng.standard_normal((n, d))
    raw_ls, raw_os, noise = rng.standard_normal(d), rng.standard_normal(()), np.asarray(0.1)
    v, lam = rng.standard_normal(n), rng.standard_normal(n)
    res = {}
    tc = bl.operators.GramOperator(X, kind=kind, path="tensor")
    alu = bl.operators.GramOperator(X, kind=kind, path="alu")
    tc.bind((raw_ls, raw_os, noise), np.float32)
    alu.bind((raw_ls, raw_os, noise), np.float32)
    # the contraction itself: -s2/2 for the first tile against float64
    orc = oops.GramOperator(X, kind=kind)
    Xs, ls, fac = orc._scaled(raw_ls)
    Xs32 = (fac * X / oops.softplus(raw_ls)).astype(np.float32).astype(np.float64)
    acc = tc.tile_distances(0, 0)
    m, w = min(n, 128), min(n, 256)
    ref = Xs32[:m] @ Xs32[:w].T - 0.5 * (Xs32[:w] ** 2).sum(-1)[None, :]
    res["tile_abs_err"] = float(np.abs(acc[:m, :w] - ref).max())
    res["tile_scale"] = float(np.abs(ref).max())
    y64 = orc.matvec(v, raw_ls, raw_os, noise)
    y_tc = tc.matvec(bl.asarray(v.astype(np.float32))).numpy()
    y_alu = alu.matvec(bl.asarray(v.astype(np.float32))).numpy()
    res["matvec_tc_vs_f64"], res["matvec_alu_vs_f64"] = rel(y_tc, y64), rel(y_alu, y64)
    z64, g64 = orc.vjp(v, lam, raw_ls, raw_os, noise)
    for name, op in (("tc", tc), ("alu", alu)):
        op.grad_zero(np.float32)
        z = op.vjp(bl.asarray(v.astype(np.float32)), bl.asarray(lam.astype(np.float32))).numpy()
        g = [x.numpy() for x in op.grad_export(np.float32)]
        res[f"vjp_z_{name}_vs_f64"] = rel(z, z64)
        res[f"vjp_dls_{name}_vs_f64"] = rel(g[0], g64[0])
        res[f"vjp_dos_{name}_vs_f64"] = rel(g[1], g64[1])
    out[f"{kind}_n{n}_d{d}"] = res
    print(kind, n, d, res, flush=True)

N, d = 45000, 9
X = rng.standard_normal((N, d))
for kind in ("matern32", "rbf"):
    for path in ("tensor", "alu"):
        op = bl.operators.GramOperator(X, kind=kind, path=path)
        v = bl.asarray(rng.standard_normal(N).astype(np.float32))
        lam = bl.asarray(rng.standard_normal(N).astype(np.float32))
        op.bind((rng.standard_normal(d), rng.standard_normal(()), np.asarray(0.1)), np.float32)
        y = bl.empty((N,), np.float32)
        ms_mv = timed(lambda: op.matvec(v, out=y))
        op.grad_zero(np.float32)
        ms_vjp = timed(lambda: op.vjp(v, lam))
        out[f"time_{kind}_{path}"] = {"matvec_ms": ms_mv, "vjp_ms": ms_vjp}
        print(kind, path, out[f"time_{kind}_{path}"], flush=True)
print(json.dumps(out))
