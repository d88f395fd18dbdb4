"""The solver oracle (`oracle/solvers.py`: CG / PCG, partial Cholesky, preconditioner, GP log-marginal
likelihood) against the golden vectors produced by the reference's own `cg.py`, `low_rank.py` and
`util/gp_util.py` (`oracle/make_golden_solvers.py`).  CPU only."""

import numpy as np
import pytest
from conftest import golden, golden_names, rel_err

from oracle import solvers


def dt(g):
    return np.float64 if bool(g["x64"]) else np.float32


@pytest.mark.parametrize("name", golden_names("cg_"))
def test_cg_matches_reference(name):
    g = golden(name)
    d = dt(g)
    A, b = g["A"].astype(d), g["b"].astype(d)
    tol = 1e-10 if d == np.float64 else 2e-4
    for steps in g["steps"]:
        x, r = solvers.pcg_fixed_step(lambda v: A @ v, b, lambda v: v, int(steps))
        assert x.dtype == d
        assert rel_err(x, g[f"x_fixed_{steps}"]) < tol
        # the residual of a converged solve is rounding noise: compare it on the scale of b
        assert np.linalg.norm(r - g[f"r_fixed_{steps}"]) < tol * np.linalg.norm(b)
    x, r, nsteps = solvers.pcg_adaptive(lambda v: A @ v, b, lambda v: v, atol=float(g["atol"]), rtol=float(g["rtol"]),
                                        maxiter=int(g["maxiter"]), miniter=int(g["miniter"]))  # fmt: skip
    assert nsteps == int(g["num_steps"])
    assert rel_err(x, g["x_adaptive"]) < tol


@pytest.mark.parametrize("name", golden_names("lowrank_dense_"))
def test_partial_cholesky_and_preconditioner_match_reference(name):
    g = golden(name)
    d = dt(g)
    cov, n, rank = g["cov"].astype(d), int(g["n"]), int(g["rank"])
    tol = 1e-9 if d == np.float64 else 5e-3  # ill-conditioned on purpose (2^-6 .. 2^5)
    col, diag = (lambda i: cov[:, i]), (lambda: np.diag(cov).copy())
    assert rel_err(solvers.cholesky_partial(col, diag, n, rank), g["L_plain"]) < tol
    L, success = solvers.cholesky_partial_pivot(col, diag, n, rank)
    assert success == bool(g["success"])
    assert rel_err(L, g["L_pivot"]) < tol
    b, small = g["b"].astype(d), d(g["small"])
    assert rel_err(solvers.preconditioner_solve(L, b, small), g["P_b"]) < tol
    x, r = solvers.pcg_fixed_step(lambda v: cov @ v + small * v, b,
                                  lambda v: solvers.preconditioner_solve(L, v, small), int(g["pcg_steps"]))  # fmt: skip
    assert rel_err(x, g["x_pcg"]) < (1e-7 if d == np.float64 else 5e-2)


@pytest.mark.parametrize("name", golden_names("lowrank_gram_"))
def test_pivoted_cholesky_of_gram_matrix_matches_reference(name):
    from oracle import operators

    g = golden(name)
    d = dt(g)
    X, n, rank = g["X"].astype(d), len(g["X"]), int(g["rank"])
    op = operators.GramOperator(X, kind="matern32")
    raw_ls, raw_os = g["raw_lengthscale"].astype(d), g["raw_outputscale"].astype(d)
    Kmat = np.stack([op.matvec(e, raw_ls, raw_os, d(0)) for e in np.eye(n, dtype=d)], axis=1)
    L, success = solvers.cholesky_partial_pivot(lambda i: Kmat[:, i], lambda: np.diag(Kmat).copy(), n, rank)
    tol = 1e-10 if d == np.float64 else 1e-4
    assert success == bool(g["success"])
    assert rel_err(L, g["L"]) < tol
    assert rel_err(solvers.preconditioner_solve(L, g["v"].astype(d), d(g["noise"])), g["P_v"]) < tol


@pytest.mark.parametrize("name", golden_names("logml_"))
def test_gp_log_marginal_likelihood_value_and_gradient_match_reference(name):
    g = golden(name)
    model = solvers.GPLogML(g["X"], g["y"], kind=str(g["kind"]), krylov_depth=int(g["K"]), probes=g["probes"],
                            rank=int(g["rank"]), cg_steps=int(g["cg_steps"]), noise_min=float(g["noise_min"]))  # fmt: skip
    value, (d_ls, d_os, d_noise, d_const) = model.value_and_grad(
        g["raw_lengthscale"], g["raw_outputscale"], g["raw_noise"], g["constant_value"])  # fmt: skip
    assert rel_err(value, g["value"]) < 1e-10
    assert rel_err(d_ls, g["d_raw_lengthscale"]) < 1e-8
    assert rel_err(d_os, g["d_raw_outputscale"]) < 1e-8
    assert rel_err(d_noise, g["d_raw_noise"]) < 1e-8
    assert rel_err(d_const, g["d_constant_value"]) < 1e-8
