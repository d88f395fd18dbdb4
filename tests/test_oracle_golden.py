"""Pin the NumPy oracle against golden vectors produced by the reference's own sources
(`oracle/make_golden.py`).  Tolerances: fp64 1e-10 relative, fp32 goldens 1e-4/1e-5
(BASELINE.json north_star)."""

import os

import numpy as np
import pytest
from conftest import golden, golden_names, rel_err

from oracle import krylov, operators

OPS = {"dense": operators.DenseOperator, "sym": operators.SymDenseOperator}


def tol(g, grad=False):
    if bool(g["x64"]):
        return 1e-10
    return 2e-4 if grad else 2e-5


def cast(g, *names):
    dt = np.float64 if bool(g["x64"]) else np.float32
    return [np.asarray(g[n], dtype=dt) for n in names]


@pytest.mark.parametrize("name", golden_names("arnoldi_"))
def test_arnoldi_forward_and_adjoint_match_reference(name):
    g = golden(name)
    A, v, dQ, dH, dr, dc = cast(g, "A", "v", "dQ", "dH", "dr", "dc")
    alg = krylov.Hessenberg(
        OPS[str(g["matvec"])](), int(g["K"]), reortho=str(g["reortho"]),
        reortho_vjp=str(g["reortho_vjp"]),
    )  # fmt: skip
    (Q, H, r, c), pull = alg.vjp(v, A)
    assert Q.dtype == A.dtype
    # Hilbert matrices (cond ~ 1e13 at n=10): rounding-order differences between BLAS
    # back-ends are amplified by the conditioning; the reference itself only checks
    # identities to sqrt(eps) there (test_hessenberg_forward.py:58-66)
    amp = 1e5 if "hilbert" in name else 1.0
    if "hilbert" in name and not bool(g["x64"]):
        # fp32 + Hilbert: H[4,3] ~ 1e-4 (near breakdown); only the identities of the
        # reference's own test hold (test_hessenberg_forward.py:58-66)
        small = np.sqrt(np.finfo(np.float32).eps)
        eK = np.eye(int(g["K"]))[-1]
        assert np.allclose(A @ Q - Q @ H - np.outer(r, eK), 0.0, atol=small)
        assert np.allclose(Q.T @ Q, np.eye(int(g["K"])), atol=small)
        assert np.allclose(Q[:, 0], c * v, atol=small)
        return
    for mine, ref in ((Q, "Q"), (H, "H"), (c, "c")):
        assert rel_err(mine, g[ref]) < amp * tol(g), ref
    if int(g["K"]) == len(v):  # full rank: the residual is rounding noise
        assert np.linalg.norm(r) < 1e-12 * np.linalg.norm(A)
    else:
        assert rel_err(r, g["r"]) < amp * tol(g)
    dv, dp = pull((dQ, dH, dr, dc))
    assert rel_err(dv, g["dv_adjoint"]) < amp * tol(g, True)
    assert rel_err(dp, g["dp_adjoint"]) < amp * tol(g, True)
    # the reference's own test: adjoint == autodiff up to 10 sqrt(eps)
    # (/root/reference/tests/test_arnoldi/test_hessenberg_adjoint.py:41-46)
    if str(g["reortho"]) == "full" or int(g["K"]) == 1:
        small = 10 * np.sqrt(np.finfo(A.dtype).eps)
        if "hilbert" in name:  # gradient entries up to 1e9: compare in norm
            assert rel_err(dv, g["dv_autodiff"]) < 1e-5
            assert rel_err(dp, g["dp_autodiff"]) < 1e-5
        else:
            assert np.allclose(dv, g["dv_autodiff"], atol=small, rtol=small)
            assert np.allclose(dp, g["dp_autodiff"], atol=small, rtol=small)


@pytest.mark.parametrize("name", golden_names("tridiag_"))
def test_tridiag_forward_and_adjoint_match_reference(name):
    g = golden(name)
    A, v = cast(g, "A", "v")
    alg = krylov.tridiag(OPS[str(g["matvec"])](), int(g["K"]), reortho=str(g["reortho"]))
    ((Qt, (alpha, beta)), (q_rem, b_rem)), pull = alg.vjp(v, A)
    for mine, ref in ((Qt, "Qt"), (alpha, "alpha"), (beta, "beta")):
        assert rel_err(mine, g[ref]) < tol(g), ref
    if int(g["K"]) == len(v):
        return  # full rank: remainder is rounding noise and its cotangent divides by it
    assert rel_err(q_rem, g["q_rem"]) < tol(g)
    assert rel_err(b_rem, g["b_rem"]) < tol(g)
    dQt, dalpha, dbeta, dq_rem, db_rem = cast(g, "dQt", "dalpha", "dbeta", "dq_rem", "db_rem")
    dv, dp = pull(((dQt, (dalpha, dbeta)), (dq_rem, db_rem)))
    assert rel_err(dv, g["dv_adjoint"]) < tol(g, True)
    assert rel_err(dp, g["dp_adjoint"]) < tol(g, True)
    # the reference's own test (test_tridiag_adjoint.py:46-50) uses the symmetrised matvec;
    # for `p @ s` the three-term adjoint returns the gradient of the symmetric problem
    if str(g["matvec"]) == "sym" or str(g["reortho"]) == "full":
        assert np.allclose(dv, g["dv_autodiff"], atol=1e-4, rtol=1e-4)
        assert np.allclose(dp, g["dp_autodiff"], atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("name", golden_names("sparse_coo_"))
def test_sparse_operand_matches_reference(name):
    g = golden(name)
    data, v = cast(g, "data", "v")
    n = int(g["n"])
    for op in (operators.CooOperator(g["row"], g["col"], (n, n)),
               operators.CsrFastOperator(g["row"], g["col"], (n, n))):  # fmt: skip
        alg = krylov.tridiag(op, int(g["K"]), reortho=str(g["reortho"]))
        ((Qt, (alpha, beta)), (q_rem, b_rem)), pull = alg.vjp(v, data)
        assert rel_err(alpha, g["alpha"]) < tol(g)
        assert rel_err(beta, g["beta"]) < tol(g)
        assert rel_err(Qt, g["Qt"]) < 10 * tol(g)
        dQt, dalpha, dbeta, dq_rem, db_rem = cast(g, "dQt", "dalpha", "dbeta", "dq_rem", "db_rem")
        dv, dp = pull(((dQt, (dalpha, dbeta)), (dq_rem, db_rem)))
        assert rel_err(dv, g["dv"]) < 5 * tol(g, True)
        assert rel_err(dp, g["dp"]) < 5 * tol(g, True)
        z = np.zeros_like
        dv0, dp0 = pull(((z(dQt), (dalpha, dbeta)), (z(dq_rem), z(db_rem))))
        assert rel_err(dv0, g["dv_slqcot"]) < 5 * tol(g, True)
        assert rel_err(dp0, g["dp_slqcot"]) < 5 * tol(g, True)


@pytest.mark.parametrize("name", golden_names("slq_"))
def test_slq_value_and_grad_match_reference(name):
    g = golden(name)
    A, probes = cast(g, "A", "probes")
    op = OPS[str(g["matvec"])]()
    integrand = krylov.IntegrandSPD(np.log, lambda x: 1.0 / x, int(g["K"]), op)
    vals = np.array([integrand(p, A) for p in probes])
    assert rel_err(vals, g["probe_values_adjoint"]) < tol(g)
    value, (grad,) = krylov.hutchinson_value_and_grad(integrand, probes, A)
    assert rel_err(value, g["value_adjoint"]) < tol(g)
    assert rel_err(grad, g["grad_adjoint"]) < tol(g, True)
    assert rel_err(grad, g["grad_autodiff"]) < 10 * tol(g, True)
    assert rel_err(krylov.hutchinson_mean(integrand, probes, A), g["value_adjoint"]) < tol(g)
    dv0 = np.stack([integrand.value_and_grad(p, A)[1][0] for p in probes])
    assert rel_err(dv0, g["probe_dv0_adjoint"]) < tol(g, True)
    reuse = krylov.IntegrandSPDReuse(np.log, lambda x: 1.0 / x, int(g["K"]), op)
    value_r, (grad_r,) = krylov.hutchinson_value_and_grad(reuse, probes, A)
    assert rel_err(value_r, g["value_reuse"]) < tol(g)
    assert rel_err(grad_r, g["grad_reuse"]) < tol(g, True)


@pytest.mark.parametrize("name", golden_names("gp_kernels_"))
@pytest.mark.parametrize("kind", ["matern32", "matern12", "rbf"])
def test_gp_kernels_match_reference(name, kind):
    g = golden(name)
    X, v, lam, raw_ls, raw_os = cast(g, "X", "v", "lam", "raw_lengthscale", "raw_outputscale")
    op = operators.GramOperator(X, kind=kind, block=16)
    noise = X.dtype.type(0.0)
    y = op.matvec(v, raw_ls, raw_os, noise)
    # Matern-1/2 evaluates exp(-sqrt(s2 + eps)) at s2 = rounding noise on the diagonal:
    # the value itself is only defined to ~sqrt(eps) there (any summation order)
    amp = 1000.0 if kind == "matern12" else 5.0
    assert rel_err(y, g[f"{kind}_y"]) < amp * tol(g)
    xbar, (dls, dos, dnoise) = op.vjp(v, lam, raw_ls, raw_os, noise)
    assert rel_err(xbar, g[f"{kind}_dv"]) < amp * tol(g)
    if kind != "matern12":  # d/ds2 at the noisy diagonal is O(1/sqrt(eps)) * noise
        assert rel_err(dls, g[f"{kind}_dls"]) < amp * tol(g, True)
    assert rel_err(dos, g[f"{kind}_dos"]) < amp * tol(g, True)
    assert np.isclose(dnoise, lam @ v)
    assert rel_err(operators.softplus(g["softplus_x"]), g["softplus_y"]) < 1e-6


def test_wave_stencil_and_expm_action_match_reference():
    g = golden("pde_wave_g8_k6_f64")
    grid, K = int(g["g"]), int(g["K"])
    op = operators.WaveStencilOperator(grid, g["stencil"])
    np.testing.assert_allclose(g["stencil"], op.stencil_laplacian(float(g["dx"])), rtol=1e-12)
    y0, lam, scale = g["y0"].ravel(), g["lam"].ravel(), g["scale"]
    assert rel_err(op.matvec(y0, scale), g["rhs"].ravel()) < 1e-12
    xbar, (dscale,) = op.vjp(y0, lam, scale)
    assert rel_err(xbar, g["rhs_dx"].ravel()) < 1e-12
    assert rel_err(dscale, g["rhs_dscale"]) < 1e-12
    out = krylov.expm_action(op, K, float(g["t1"]), y0, scale)
    assert rel_err(out, g["expm_out"].ravel()) < 1e-10
    dy0, dsc = krylov.expm_action_vjp(op, K, float(g["t1"]), y0, (scale,), g["u"].ravel())
    assert rel_err(dy0, g["loss_dy0"].ravel()) < 1e-9
    assert rel_err(dsc, g["loss_dscale"]) < 1e-9


@pytest.mark.parametrize("krylov_depth", [1, 5, 10])
def test_oracle_arnoldi_forward_supports_complex(krylov_depth, nrows=10):
    """/root/reference/tests/test_arnoldi/test_hessenberg_forward.py:10-37 with dtype=complex: the
    forward pass conjugates (`arnoldi.py:66,87,92,95`); the CUDA path is real-only."""
    rng = np.random.default_rng(1)
    A = rng.standard_normal((nrows, nrows)) + 1j * rng.standard_normal((nrows, nrows))
    v = rng.standard_normal(nrows) + 1j * rng.standard_normal(nrows)
    Q, H, r, c = krylov.arnoldi_forward(operators.DenseOperator(), krylov_depth, v, A)
    small = np.sqrt(np.finfo(np.float64).eps)
    e0, ek = np.eye(krylov_depth)[[0, -1], :]
    assert np.allclose(A @ Q - Q @ H - np.outer(r, ek), 0.0, atol=small)
    assert np.allclose(Q.T.conj() @ Q - np.eye(krylov_depth), 0.0, atol=small)
    assert np.allclose(Q @ e0, c * v, atol=small)


def test_oracle_error_conventions():
    # arnoldi.py:16-19 (TypeError), :58-60 (ValueError "depth"), lanczos.py:148-149 (ValueError)
    with pytest.raises(TypeError, match="Unexpected input"):
        krylov.Hessenberg(operators.DenseOperator(), 1, reortho="None")
    for depth in (0, 3):
        with pytest.raises(ValueError, match="depth"):
            krylov.Hessenberg(operators.DenseOperator(), depth, reortho="none")(np.ones(2), np.eye(2))
    with pytest.raises(ValueError, match="unsupported"):
        krylov.tridiag(operators.DenseOperator(), 1, reortho="partial")


# ---- the library's symmetric loops, restated in the oracle, against the reference restatement ---------
def _symmetric_tridiag_vjp(op, K, v, params, cot_alpha_beta, dr=None, dQ=None):
    """`tridiag(reortho="full")` through `arnoldi_forward(symmetric=True)` and the banded adjoint."""
    Q, H, r, c = krylov.arnoldi_forward(op, K, v, *params, symmetric=True)
    dalpha, dbeta = cot_alpha_beta
    dH = np.diag(dalpha) + 0.5 * (np.diag(dbeta, 1) + np.diag(dbeta, -1))
    n = len(v)
    dv, dp = krylov.arnoldi_adjoint(
        op, params, Q=Q, H=H, r=r, c=c, dQ=np.zeros((n, K)) if dQ is None else dQ, dH=dH,
        dr=np.zeros(n) if dr is None else dr, dc=0.0, reortho="full", symmetric=True, tridiagonal_cotangent=True,
    )  # fmt: skip
    return (Q, H, r, c), (dv, dp)


def test_symmetric_loops_match_the_reference_loops_on_a_sparse_spd_operand():
    """Why the shortcuts are legal, checked in float64 without a GPU: with a symmetric operand and full
    re-orthogonalisation `H` is tridiagonal up to rounding, and so is everything the three flag bits drop."""
    rng = np.random.default_rng(0)
    n, K = 600, 40
    offs = [1, 7, 31]
    row = np.concatenate([np.arange(n)] + [np.arange(o, n) for o in offs] + [np.arange(0, n - o) for o in offs])
    col = np.concatenate([np.arange(n)] + [np.arange(0, n - o) for o in offs] + [np.arange(o, n) for o in offs])
    vals = [-rng.uniform(0, 1, n - o) for o in offs]
    data = np.concatenate([np.full(n, 8.0)] + vals + vals)
    op = operators.CsrFastOperator(row.astype(np.int32), col.astype(np.int32), (n, n))
    v = rng.standard_normal(n)
    dalpha, dbeta, dr = rng.standard_normal(K), rng.standard_normal(K - 1), rng.standard_normal(n)

    Q0, H0, r0, c0 = krylov.arnoldi_forward(op, K, v, data)
    assert np.abs(np.triu(H0, 2)).max() < 1e-13 * np.abs(H0).max()
    dH = np.diag(dalpha) + 0.5 * (np.diag(dbeta, 1) + np.diag(dbeta, -1))
    for with_dr in (False, True):
        dr_ = dr if with_dr else np.zeros(n)
        dv0, dp0 = krylov.arnoldi_adjoint(op, (data,), Q=Q0, H=H0, r=r0, c=c0, dQ=np.zeros((n, K)), dH=dH, dr=dr_,
                                          dc=0.0, reortho="full")  # fmt: skip
        (Q1, H1, r1, c1), (dv1, dp1) = _symmetric_tridiag_vjp(op, K, v, (data,), (dalpha, dbeta), dr=dr_)
        # the local first pass skips rows j < i-1; the second pass's coefficients complete those entries of H
        assert np.abs(np.triu(H1, 2) - np.triu(H0, 2)).max() < 1e-13 * np.abs(H0).max()
        assert rel_err(np.diag(H1), np.diag(H0)) < 1e-13 and rel_err(np.diag(H1, 1), np.diag(H0, 1)) < 1e-13
        assert rel_err(Q1, Q0) < 1e-12 and rel_err(r1, r0) < 1e-11
        assert np.abs(Q1.T @ Q1 - np.eye(K)).max() < 1e-14
        assert rel_err(dv1, dv0) < 1e-11 and rel_err(dp1[0], dp0[0]) < 1e-11
    # a dense dQ keeps the general Gamma (only `Lambda beta_plus` is shortened)
    dQ = rng.standard_normal((n, K))
    dv0, dp0 = krylov.arnoldi_adjoint(op, (data,), Q=Q0, H=H0, r=r0, c=c0, dQ=dQ, dH=dH, dr=dr, dc=0.3, reortho="full")
    dv1, dp1 = krylov.arnoldi_adjoint(op, (data,), Q=Q0, H=H0, r=r0, c=c0, dQ=dQ, dH=dH, dr=dr, dc=0.3, reortho="full",
                                      symmetric=True, tridiagonal_cotangent=True)  # fmt: skip
    assert rel_err(dv1, dv0) < 1e-11 and rel_err(dp1[0], dp0[0]) < 1e-11


@pytest.mark.parametrize("name", [n for n in golden_names("tridiag_") if "_full_" in n and "k12" not in n])
def test_symmetric_loops_reproduce_the_reference_goldens(name):
    """The same restatement against the outputs of the reference's own sources (tests/golden)."""
    g = golden(name)
    A, v = cast(g, "A", "v")
    K = int(g["K"])
    dQt, dalpha, dbeta, dq_rem, db_rem = cast(g, "dQt", "dalpha", "dbeta", "dq_rem", "db_rem")
    op = OPS[str(g["matvec"])]()
    Q, H, r, c = krylov.arnoldi_forward(op, K, v, A, symmetric=True)
    T = 0.5 * (H + H.T)
    assert rel_err(np.diag(T), g["alpha"]) < tol(g) and rel_err(np.diag(T, 1), g["beta"]) < tol(g)
    assert rel_err(Q.T, g["Qt"]) < tol(g)
    # cotangent of (r/||r||, ||r||) -> dr (lanczos.py:166), then the symmetric adjoint
    norm = np.linalg.norm(r)
    dr = dq_rem / norm + (db_rem / norm - np.dot(r, dq_rem) / norm**3) * r
    dH = np.diag(dalpha) + 0.5 * (np.diag(dbeta, 1) + np.diag(dbeta, -1))
    dv, dp = krylov.arnoldi_adjoint(op, (A,), Q=Q, H=H, r=r, c=c, dQ=dQt.T, dH=dH, dr=dr, dc=np.zeros((), A.dtype),
                                    reortho="full", symmetric=True, tridiagonal_cotangent=True)  # fmt: skip
    assert rel_err(dv, g["dv_adjoint"]) < tol(g, True)
    assert rel_err(dp[0], g["dp_adjoint"]) < tol(g, True)


# ---- active-column loops (headline-size parity test) against the literal restatement -----------------------------
@pytest.mark.parametrize("with_dense_cotangent", [False, True])
def test_active_column_loops_equal_the_literal_loops(with_dense_cotangent):
    rng = np.random.default_rng(3)
    n, K = 700, 23
    offs = [1, 5, 19]
    row = np.concatenate([np.arange(n)] + [np.arange(o, n) for o in offs] + [np.arange(0, n - o) for o in offs])
    col = np.concatenate([np.arange(n)] + [np.arange(0, n - o) for o in offs] + [np.arange(o, n) for o in offs])
    vals = [-rng.uniform(0, 1, n - o) for o in offs]
    data = np.concatenate([np.full(n, 8.0)] + vals + vals)
    op = operators.CsrFastOperator(row.astype(np.int32), col.astype(np.int32), (n, n))
    v = rng.standard_normal(n)
    ((Qt_r, (a_r, b_r)), (q_r, n_r)), pull_r = krylov.tridiag(op, K, reortho="full").vjp(v, data)
    ((Qt, (a, b)), (q, nrm)), pull = krylov.tridiag_full_active(op, K, v, data)
    assert rel_err(Qt, Qt_r) < 1e-13 and rel_err(a, a_r) < 1e-13 and rel_err(b, b_r) < 1e-13
    assert rel_err(q, q_r) < 1e-12 and abs(nrm - n_r) < 1e-13 * n_r
    da, db = rng.standard_normal(K), rng.standard_normal(K - 1)
    if with_dense_cotangent:
        cot = ((rng.standard_normal((K, n)), (da, db)), (rng.standard_normal(n), rng.standard_normal()))
        cot_r = cot
    else:
        cot = ((None, (da, db)), (None, None))
        cot_r = ((np.zeros_like(Qt_r), (da, db)), (np.zeros_like(q_r), np.zeros(())))
    dv, dp = pull(cot)
    dv_r, dp_r = pull_r(cot_r)
    assert rel_err(dv, dv_r) < 1e-11 and rel_err(dp, dp_r) < 1e-11


def test_symmetry_check_on_the_hessenberg_matrix():
    """`lanczos.hessenberg_is_tridiagonal`: what decides between the symmetric and the general adjoint loops."""
    from experiments_lanczos_adjoints_b200.lanczos import hessenberg_is_tridiagonal

    rng = np.random.default_rng(0)
    n, K = 80, 12
    S = rng.standard_normal((n, n))
    S = S + S.T + 20 * np.eye(n)
    v = rng.standard_normal(n)
    for dtype in (np.float32, np.float64):
        _, H, _, _ = krylov.arnoldi_forward(operators.DenseOperator(), K, v.astype(dtype), S.astype(dtype), symmetric=True)
        assert hessenberg_is_tridiagonal(H)
        N = S + 0.05 * np.triu(rng.standard_normal((n, n)), 1)  # slightly non-symmetric operand
        _, H, _, _ = krylov.arnoldi_forward(operators.DenseOperator(), K, v.astype(dtype), N.astype(dtype), symmetric=True)
        assert not hessenberg_is_tridiagonal(H)
    assert hessenberg_is_tridiagonal(np.ones((1, 1)))
    assert not hessenberg_is_tridiagonal(np.array([[1.0, np.nan], [0.5, 1.0]]))


# ---- round-2 fixtures: SuiteSparse file through suite_sparse_load, batched initial conditions ---------------------
@pytest.mark.parametrize("name", golden_names("suitesparse_"))
def test_oracle_matches_the_suitesparse_fixture(name):
    """`exp_util.suite_sparse_load("1138_bus")` -> BCOO operand -> `tridiag(reortho="full")` + VJP, made by the
    reference's own sources (oracle/make_golden_r2.py).  The float32 fixture is compared at float32 tolerances (the
    oracle runs in float64)."""
    g = golden(name)
    n, K = int(g["n"]), int(g["K"])
    ref = krylov.tridiag(operators.CooOperator(g["row"], g["col"], (n, n)), K, reortho="full")
    ((Qt, (alpha, beta)), (q_rem, b_rem)), pull = ref.vjp(g["v"], g["data"])
    t_val, t_grad = (1e-12, 1e-12) if bool(g["x64"]) else (1e-5, 1e-4)
    assert rel_err(alpha, g["alpha"]) < t_val and rel_err(beta, g["beta"]) < t_val and rel_err(Qt, g["Qt"]) < 10 * t_val
    dv, dp = pull(((g["dQt"], (g["dalpha"], g["dbeta"])), (g["dq_rem"], g["db_rem"])))
    assert rel_err(dv, g["dv"]) < t_grad and rel_err(dp, g["dp"]) < t_grad
    z = np.zeros_like
    dv0, dp0 = pull(((z(Qt), (g["dalpha"], g["dbeta"])), (z(q_rem), z(b_rem))))
    assert rel_err(dv0, g["dv_slqcot"]) < t_grad and rel_err(dp0, g["dp_slqcot"]) < t_grad


def test_matrix_market_loader_reproduces_suite_sparse_load_order():
    """`SparseOperator.from_matrix_market` must hand out the COO entries -- and so the parameter vector and its
    gradient -- in the order `scipy.io.mmread` gives `suite_sparse_load` (stored lower-triangular entries in file
    order, then the mirrored strict upper triangle; exp_util.py:35-42): index work, bit-exact."""
    from conftest import GOLDEN_DIR as GOLDEN

    import experiments_lanczos_adjoints_b200 as bl

    g = golden("suitesparse_1138_bus_k20_f64")
    op, data = bl.operators.SparseOperator.from_matrix_market(os.path.join(GOLDEN, "1138_bus.mtx"))
    assert op.shape == (1138, 1138) and len(data) == 4054
    assert np.array_equal(op._coo[0], g["row"]) and np.array_equal(op._coo[1], g["col"])
    assert np.array_equal(data, g["data"])


def test_oracle_matches_the_batched_initial_conditions_fixture():
    """`jax.vmap(solve, in_axes=(0, None))(y0s, scale)` (train.py:104-110): per-run outputs and `dy0`, the parameter
    cotangent summed over the batch."""
    g = golden("pde_wave_batch_g8_k6_f64")
    gg, K, B, t1 = int(g["g"]), int(g["K"]), int(g["B"]), float(g["t1"])
    op = operators.WaveStencilOperator(gg, g["stencil"])
    dscale = 0.0
    for b in range(B):
        y0 = g["y0s"][b].ravel()
        assert rel_err(krylov.expm_action(op, K, t1, y0, g["scale"]), g["expm_out"][b].ravel()) < 1e-12
        dy0, ds = krylov.expm_action_vjp(op, K, t1, y0, (g["scale"],), g["u"][b].ravel())
        assert rel_err(dy0, g["loss_dy0s"][b].ravel()) < 1e-11
        dscale = dscale + ds
    assert rel_err(dscale, g["loss_dscale"]) < 1e-11
