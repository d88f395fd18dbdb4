"""One Gram matvec + one VJP at the UCI-protein shape (ncu target).  KIND / DTYPE / GP_N / GRAM_PATH from the environment."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl

N, d = int(os.environ.get("GP_N", 45000)), 9
kind = os.environ.get("KIND", "matern32")
dtype = np.float32 if os.environ.get("DTYPE", "f32") == "f32" else np.float64
rng = np.random.default_rng(0)
X = rng.standard_normal((N, d))
op = bl.operators.GramOperator(X, kind=kind, path=os.environ.get("GRAM_PATH", "auto"))
v = bl.asarray(rng.standard_normal(N).astype(dtype))
lam = bl.asarray(rng.standard_normal(N).astype(dtype))
op.bind((rng.standard_normal(d), rng.standard_normal(()), np.asarray(0.1)), dtype)
y = bl.empty((N,), dtype)
for _ in range(int(os.environ.get("REPS", 2))):
    op.matvec(v, out=y)
    op.grad_zero(dtype)
    op.vjp(v, lam)
bl.synchronize()
print("ok", float(np.abs(y.numpy()).sum()))
