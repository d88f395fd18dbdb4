"""GPU parity tests of the solver half of the GP path: CG / PCG, partial (pivoted)
Cholesky, the low-rank preconditioner and the GP log-marginal likelihood with its gradient -- through the
host layer -> C ABI -> `csrc/solve.cu`, against the golden vectors produced by the reference's own
`cg.py` / `low_rank.py` / `util/gp_util.py` and against the NumPy oracle on seeded problems.
Tolerances: fp64 1e-10 (1e-8 where a converged CG residual amplifies rounding), fp32 1e-5 / 1e-4."""

import numpy as np
import pytest
from conftest import golden, golden_names, rel_err

import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import cg, gp, low_rank
from oracle import operators as oops
from oracle import solvers

pytestmark = pytest.mark.gpu


def dt(g):
    return np.float64 if bool(g["x64"]) else np.float32


def dense(A, d):
    return bl.operators.bound(bl.operators.DenseOperator(len(A)), A.astype(d))


@pytest.mark.parametrize("name", golden_names("cg_"))
def test_cg_matches_reference_golden(name):
    g = golden(name)
    d = dt(g)
    A, b = dense(g["A"], d), g["b"].astype(d)
    tol = 1e-10 if d == np.float64 else 2e-4
    for steps in g["steps"]:
        x, info = cg.cg_fixed_step(int(steps))(A, b)
        assert rel_err(x.numpy(), g[f"x_fixed_{steps}"]) < tol
        assert np.linalg.norm(info["residual_abs"].numpy() - g[f"r_fixed_{steps}"]) < tol * np.linalg.norm(b)
        assert info["residual_rel"].shape == b.shape
    solve = cg.cg_adaptive(atol=float(g["atol"]), rtol=float(g["rtol"]), maxiter=int(g["maxiter"]), miniter=int(g["miniter"]),
                           check_every=3)  # fmt: skip
    x, info = solve(A, b)
    assert info["num_steps"] == int(g["num_steps"])  # frozen at the iteration where the reference's loop stops
    assert rel_err(x.numpy(), g["x_adaptive"]) < tol


def test_reference_cg_scenarios():
    """tests/test_cg/test_cg.py:10-31, 86-100: solves an SPD system; more matvecs shrink the residual."""
    rng = np.random.default_rng(0)
    U, _ = np.linalg.qr(rng.standard_normal((9, 9)))
    A = (U * np.arange(1.0, 10.0)) @ U.T
    b = np.arange(1.0, 10.0)
    x, _ = cg.cg_fixed_step(9)(dense(A, np.float64), b)
    assert np.allclose(x.numpy(), np.linalg.solve(A, b))
    x, _ = cg.cg_adaptive(atol=1e-5, rtol=1e-5, maxiter=100, miniter=1)(dense(A, np.float64), b)
    assert np.allclose(x.numpy(), np.linalg.solve(A, b))
    error = 100.0
    for n in range(9):
        _, info = cg.cg_fixed_step(n)(dense(A, np.float64), b)
        now = np.linalg.norm(info["residual_abs"].numpy())
        assert now < error
        error = now


def test_cg_on_sparse_operand_and_linear_solve_vjp():
    """PCG on the SELL operand (n = 20 000) against the oracle, and the `custom_linear_solve` rule:
    d<c, x>/db = A^{-1} c and d<c, x>/dtheta = -(A^{-1} c)_row x_col per stored entry."""
    from experiments_lanczos_adjoints_b200 import synthetic

    n = 20_000
    row, col, data = synthetic.banded_spd_coo(n, bands=4, seed=3)
    rng = np.random.default_rng(1)
    b, c = rng.standard_normal(n), rng.standard_normal(n)
    op = bl.operators.SparseOperator(row, col, (n, n))
    A = bl.operators.bound(op, data)
    solve = cg.cg_fixed_step(60)
    (x, info), pull = solve.vjp(A, b)
    import scipy.sparse

    M = scipy.sparse.coo_matrix((data, (row, col)), shape=(n, n)).tocsr()
    xo, _ = solvers.pcg_fixed_step(lambda v: M @ v, b, lambda v: v, 60)
    assert rel_err(x.numpy(), xo) < 1e-10
    (dtheta,), db = pull(c)
    lam = scipy.sparse.linalg.spsolve(M.tocsc(), c)
    assert rel_err(db.numpy(), lam) < 1e-8
    assert rel_err(dtheta.numpy(), -lam[row] * xo[col]) < 1e-8


@pytest.mark.parametrize("name", golden_names("lowrank_dense_"))
def test_partial_cholesky_and_preconditioner_match_reference_golden(name):
    g = golden(name)
    d = dt(g)
    n, rank = int(g["n"]), int(g["rank"])
    tol = 1e-9 if d == np.float64 else 5e-3  # ill-conditioned on purpose
    lazy = dense(g["cov"], d)
    L, info = low_rank.cholesky_partial(rank=rank)(lazy, n)
    assert L.shape == (n, rank) and info == {}
    assert rel_err(L.numpy(), g["L_plain"]) < tol
    Lp, info = low_rank.cholesky_partial_pivot(rank=rank)(lazy, n)
    assert info["success"] == bool(g["success"])
    assert rel_err(Lp.numpy(), g["L_pivot"]) < tol
    pre, _ = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=rank))(lazy, n)
    b, small = g["b"].astype(d), float(g["small"])
    assert rel_err(pre(b, small).numpy(), g["P_b"]) < tol
    shifted = dense(g["cov"] + small * np.eye(n), d)
    x, _ = cg.pcg_fixed_step(int(g["pcg_steps"]))(shifted, b, pre.bind(small))
    assert rel_err(x.numpy(), g["x_pcg"]) < (1e-7 if d == np.float64 else 5e-2)


def test_reference_low_rank_scenarios():
    """tests/test_low_rank/test_low_rank.py: full rank reconstructs, shapes, pivoting improves, errors."""
    rng = np.random.default_rng(2)
    n = 10
    U, _ = np.linalg.qr(rng.standard_normal((n, n)))
    cov = (U * (0.1 + rng.uniform(size=n))) @ U.T
    lazy = dense(cov, np.float64)
    for factory in (low_rank.cholesky_partial, low_rank.cholesky_partial_pivot):
        L, _ = factory(rank=n)(lazy, n)
        assert np.allclose(L.numpy() @ L.numpy().T, cov, atol=1e-12)
        L, _ = factory(rank=4)(lazy, n)
        assert L.shape == (n, 4)
    plain, _ = low_rank.cholesky_partial(rank=n)(lazy, n)
    assert np.allclose(plain.numpy(), np.linalg.cholesky(cov), atol=1e-10)
    nopivot, _ = low_rank.cholesky_partial(rank=5)(lazy, n)
    pivot, _ = low_rank.cholesky_partial_pivot(rank=5)(lazy, n)
    assert np.linalg.norm(cov - pivot.numpy() @ pivot.numpy().T) < np.linalg.norm(cov - nopivot.numpy() @ nopivot.numpy().T)
    with pytest.raises(ValueError, match="Rank exceeds n"):
        low_rank.cholesky_partial_pivot(rank=n + 1)(lazy, n)
    with pytest.raises(ValueError, match="Rank must be positive"):
        low_rank.cholesky_partial(rank=0)(lazy, n)


@pytest.mark.parametrize("name", golden_names("lowrank_gram_"))
def test_pivoted_cholesky_of_gram_matrix_matches_reference_golden(name):
    g = golden(name)
    d = dt(g)
    n, rank = len(g["X"]), int(g["rank"])
    op = bl.operators.GramOperator(g["X"], kind="matern32")
    lazy = bl.operators.bound(op, g["raw_lengthscale"].astype(d), g["raw_outputscale"].astype(d).reshape(1), np.zeros(1, d))
    L, info = low_rank.cholesky_partial_pivot(rank=rank)(lazy, n)
    tol = 1e-10 if d == np.float64 else 1e-4
    assert info["success"] == bool(g["success"])
    assert rel_err(L.numpy(), g["L"]) < tol
    pre, _ = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=rank))(lazy, n)
    assert rel_err(pre(g["v"].astype(d), float(g["noise"])).numpy(), g["P_v"]) < tol


def test_pivoted_cholesky_at_scale_is_a_good_preconditioner():
    """n = 8192, rank 64 on a smooth Gram matrix: the pivots are distinct, the residual diagonal shrinks,
    and PCG with the preconditioner needs far fewer iterations than plain CG (the reason the reference
    uses it, gp_util.py:243-276)."""
    n, d = 8192, 4
    rng = np.random.default_rng(3)
    X = rng.uniform(size=(n, d))
    op = bl.operators.GramOperator(X, kind="matern32")
    noise = 1e-2
    params = (np.full(d, 1.5, np.float64), np.zeros(1), np.full(1, noise))
    A = bl.operators.bound(op, *params)
    L, info = low_rank.cholesky_partial_pivot(rank=64)(A, n)
    assert info["success"] and len(set(info["pivots"].tolist())) == 64
    Lh = L.numpy()
    diag = np.full(n, oops.softplus(0.0) * (1 + np.sqrt(np.finfo(np.float64).eps)) * np.exp(-np.sqrt(np.finfo(np.float64).eps)))
    resid = diag - (Lh**2).sum(1)
    assert resid.min() > -1e-10 and resid.max() < 0.5 * diag[0]
    b = rng.standard_normal(n)
    pre, _ = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=64))(A, n)
    plain = cg.cg_adaptive(atol=1e-6, rtol=0.0, maxiter=2000, miniter=1)
    precon = cg.pcg_adaptive(atol=1e-6, rtol=0.0, maxiter=2000, miniter=1)
    x0, i0 = plain(A, b)
    x1, i1 = precon(A, b, pre.bind(noise))
    assert i1["num_steps"] < 0.6 * i0["num_steps"]
    assert rel_err(x1.numpy(), x0.numpy()) < 1e-5


@pytest.mark.parametrize("name", golden_names("logml_"))
def test_gp_log_marginal_likelihood_matches_reference_golden(name):
    g = golden(name)
    X, y, d = g["X"], g["y"], g["X"].shape[1]
    probes = g["probes"]
    solve_p = cg.pcg_fixed_step(int(g["cg_steps"]))
    logdet = gp.krylov_logdet_slq(int(g["K"]), sample=lambda key: probes, num_batches=1, checkpoint=True)
    precondition = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=int(g["rank"])))
    logpdf_p = gp.logpdf_krylov_p(solve_p=solve_p, logdet=logdet)
    likelihood, p_lik = gp.likelihood_pdf_p(gp.gram_matvec(), logpdf_p, precondition=precondition,
                                            constrain=gp.constraint_greater_than(float(g["noise_min"])))  # fmt: skip
    m, p_mean = gp.mean_constant(shape_out=())
    kernels = {"matern32": gp.kernel_scaled_matern_32, "rbf": gp.kernel_scaled_rbf}
    k, p_kernel = kernels[str(g["kind"])](shape_in=(d,), shape_out=())
    assert set(p_lik) == {"raw_noise"} and set(p_mean) == {"constant_value"}
    assert set(p_kernel) == {"raw_lengthscale", "raw_outputscale"}
    loss = gp.target_logml(gp.model_gp(m, k), likelihood)
    params = dict(params_mean={"constant_value": g["constant_value"]},
                  params_kernel={"raw_lengthscale": g["raw_lengthscale"], "raw_outputscale": g["raw_outputscale"]},
                  params_likelihood={"raw_noise": g["raw_noise"]})  # fmt: skip
    value, info = loss(X, y, None, **params)
    assert rel_err(value, g["value"]) < 1e-10
    assert info["precondition"]["success"]
    (value, info), (d_mean, d_kernel, d_lik) = loss.value_and_grad(X, y, None, **params)
    assert rel_err(value, g["value"]) < 1e-10
    assert rel_err(d_kernel["raw_lengthscale"], g["d_raw_lengthscale"]) < 1e-8
    assert rel_err(d_kernel["raw_outputscale"], g["d_raw_outputscale"]) < 1e-8
    assert rel_err(d_lik["raw_noise"], g["d_raw_noise"]) < 1e-8
    assert rel_err(d_mean["constant_value"], g["d_constant_value"]) < 1e-8
    # fp32 run of the same problem: north_star tolerances (1e-5 value, 1e-4 gradient)
    value32, _ = loss(X, y.astype(np.float32), None, **params)
    assert rel_err(value32, g["value"]) < 1e-5
    (_, _), (_, d_kernel32, d_lik32) = loss.value_and_grad(X, y.astype(np.float32), None, **params)
    assert rel_err(d_kernel32["raw_lengthscale"], g["d_raw_lengthscale"]) < 1e-4
    assert rel_err(d_lik32["raw_noise"], g["d_raw_noise"]) < 1e-4
