"""BASELINE config 3: the full GP log-marginal-likelihood value + gradient (SLQ log-determinant through the
Lanczos adjoint, PCG solve with the pivoted-Cholesky preconditioner) at the UCI-protein training shape
(n = 36 560, d = 9; training defaults of run_uci.sh: depth 10, 10 probes, rank-100 preconditioner, cg_tol 1.0 ->
here 1e-2).  Prints one JSON line: wall time per evaluation and its breakdown."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import cg, gp, hutchinson, low_rank

n, d = int(os.environ.get("GP_N", 36560)), 9
K, probes_n, rank = int(os.environ.get("DEPTH", 10)), int(os.environ.get("PROBES", 10)), int(os.environ.get("RANK", 100))
dtype = np.float32 if os.environ.get("DTYPE", "f32") == "f32" else np.float64
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d))
y = (np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)).astype(dtype)
probes = (rng.integers(0, 2, size=(probes_n, n)) * 2 - 1).astype(dtype)

solve_p = cg.pcg_adaptive(rtol=0.0, atol=float(os.environ.get("CG_TOL", 1e-2)), maxiter=1000, miniter=10)
logdet = gp.krylov_logdet_slq(K, sample=lambda key: probes, num_batches=1, checkpoint=True)
precondition = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=rank))
logpdf_p = gp.logpdf_krylov_p(solve_p=solve_p, logdet=logdet)
likelihood, _ = gp.likelihood_pdf_p(gp.gram_matvec(), logpdf_p, precondition=precondition,
                                    constrain=gp.constraint_greater_than(1e-4))  # fmt: skip
m, _ = gp.mean_constant(shape_out=())
k, _ = gp.kernel_scaled_matern_32(shape_in=(d,), shape_out=())
loss = gp.target_logml(gp.model_gp(m, k), likelihood)
params = dict(params_mean={"constant_value": 0.0},
              params_kernel={"raw_lengthscale": np.full(d, 1.0), "raw_outputscale": 0.5},
              params_likelihood={"raw_noise": -1.0})  # fmt: skip


def evaluate():
    out = loss.value_and_grad(X, y, None, **params)
    bl.synchronize()
    return out


evaluate()
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    (value, info), grads = evaluate()
sec = (time.perf_counter() - t0) / reps
# breakdown
A = bl.operators.bound(likelihood.operator(X, "matern32"), np.full(d, 1.0, dtype), np.full(1, 0.5, dtype),
                       np.asarray(gp.constraint_greater_than(1e-4)(-1.0), dtype).reshape(1))  # fmt: skip
t0 = time.perf_counter()
pre, pinfo = precondition(A, n)
bl.synchronize()
t_chol = time.perf_counter() - t0
t0 = time.perf_counter()
x, sinfo = solve_p(A, y, pre.bind(float(A.params[2][0])))
bl.synchronize()
t_solve = time.perf_counter() - t0
t0 = time.perf_counter()
ld, _, _ = logdet.value_and_grad(A, None)
bl.synchronize()
t_logdet = time.perf_counter() - t0
print(json.dumps({"config": "C3 GP log-marginal likelihood value+grad", "n": n, "d": d, "dtype": np.dtype(dtype).name,
                  "krylov_depth": K, "probes": probes_n, "precond_rank": rank, "seconds_per_eval": sec,
                  "value": float(value), "cg_steps": int(sinfo["num_steps"]),
                  "breakdown_s": {"pivoted_cholesky": t_chol, "pcg_solve": t_solve, "slq_logdet_value_grad": t_logdet},
                  "grad_raw_lengthscale": np.asarray(grads[1]["raw_lengthscale"]).tolist()}))
