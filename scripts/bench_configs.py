"""Timings of the other BASELINE.json configurations (parity-test shapes, not the bench line):
C1 dense SPD n=100 K=10 fp64; C3 Gram matvec / SLQ gradient at the UCI-protein shape; C5 wave
stencil 4096^2 Arnoldi K=10 forward + adjoint.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl


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


out = {}
rng = np.random.default_rng(0)

# ---- C3: Gram matvec at N = 45 000, d = 9 (benchmark_datasets.py:141-146) ----
N, d = int(os.environ.get("GP_N", 45000)), 9
X = rng.standard_normal((N, d))
for kind in ("matern32", "rbf"):
    for dtype in (np.float32, np.float64):
        op = bl.operators.GramOperator(X, kind=kind)
        v = bl.asarray(rng.standard_normal(N).astype(dtype))
        lam = bl.asarray(rng.standard_normal(N).astype(dtype))
        op.bind((rng.standard_normal(d), rng.standard_normal(()), np.asarray(0.1)), dtype)
        y = bl.empty((N,), dtype)
        ms_mv = timed(lambda: op.matvec(v, out=y))
        op.grad_zero(dtype)
        ms_vjp = timed(lambda: op.vjp(v, lam))
        out[f"gram_{kind}_{np.dtype(dtype).name}"] = {
            "n": N, "d": d, "matvec_ms": ms_mv, "vjp_ms": ms_vjp, "pairs_per_s": N * N / (ms_mv * 1e-3)}
# SLQ log-det value + gradient, K = 10, 10 probes (the reference's training default, run_uci.sh:26)
Ntr = 36560
op = bl.operators.GramOperator(X[:Ntr], kind="matern32")
integrand = bl.lanczos.integrand_spd(np.log, 10, op)
probes = (rng.integers(0, 2, size=(10, Ntr)) * 2 - 1).astype(np.float32)
est = bl.hutchinson.hutchinson(integrand, lambda key: probes)
params = (rng.standard_normal(d).astype(np.float32), np.float32(0.3), np.float32(0.5))
t0 = time.perf_counter()
val, grads = est.value_and_grad(None, *params)
bl.synchronize()
t1 = time.perf_counter()
val, grads = est.value_and_grad(None, *params)
bl.synchronize()
out["gp_slq_logdet_grad_f32"] = {"n": Ntr, "K": 10, "probes": 10, "seconds": time.perf_counter() - t1,
                                 "first_call_seconds": t1 - t0, "value": float(val)}

# ---- C5: wave stencil, 4096^2 grid, Arnoldi K = 10 forward + adjoint (single GPU) ----
g = int(os.environ.get("PDE_G", 4096))
dx = 1.0 / (g - 1)
op = bl.operators.WaveStencilOperator(g, bl.operators.WaveStencilOperator.stencil_laplacian(dx) * dx * dx)
xs = np.linspace(0, 1, g)
y0 = np.stack([np.exp(-80 * ((xs[:, None] - 0.4) ** 2 + (xs[None, :] - 0.6) ** 2)), np.zeros((g, g))])
scale = (1.0 + 0.1 * np.sin(6 * xs)[:, None] * np.cos(4 * xs)[None, :])
for dtype in (np.float32,):
    alg = bl.arnoldi.hessenberg(op, 10, reortho="full")
    v = bl.asarray(y0.ravel().astype(dtype))
    sc = bl.asarray(scale.astype(dtype))

    def run():
        (Q, H, r, c), pull = bl.vjp(alg, v, sc)
        return pull((None, np.eye(10, dtype=dtype), None, None))

    t = timed(run, reps=3, warm=1)
    n = 2 * g * g
    out[f"wave_arnoldi_k10_{np.dtype(dtype).name}"] = {"grid": g, "n": n, "fwd_adj_ms": t}

# ---- C1: dense SPD 100 x 100, K = 10, fp64 ----
A = rng.standard_normal((100, 100))
A = A @ A.T / 100 + np.eye(100)
op = bl.operators.DenseOperator(100, sym=True)
alg = bl.lanczos.tridiag(op, 10, reortho="full")
vv = rng.standard_normal(100)
P = np.triu(A) - 0.5 * np.diag(np.diag(A))


def run_c1():
    out_, pull = bl.vjp(alg, vv, P)
    return pull(((None, (np.ones(10), np.ones(9))), (None, None)))


out["dense_n100_k10_f64_fwd_adj_ms"] = timed(run_c1, reps=5, warm=2)
print(json.dumps(out))
