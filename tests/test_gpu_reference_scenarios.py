"""The reference's own test scenarios, replayed on the CUDA path through the mirrored API.
Each test names the reference test it restates; inputs are seeded NumPy (JAX's PRNG stream is
not reproducible here), the assertions and tolerances are the reference's."""

import numpy as np
import pytest

import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import arnoldi, hutchinson, lanczos

pytestmark = pytest.mark.gpu


def symmetric_matrix_from_eigenvalues(eigvals, seed=0):
    """Stand-in for `matfree.test_util.symmetric_matrix_from_eigenvalues`: Q diag(e) Q^T."""
    n = len(eigvals)
    Q, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return (Q * np.asarray(eigvals)) @ Q.T


def dense_tridiag(d, o):
    return np.diag(d) + np.diag(o, 1) + np.diag(o, -1)


@pytest.mark.parametrize("reortho", ["full", "none"])
def test_full_rank_reconstruction_is_exact(reortho, ndim=12):
    # /root/reference/tests/test_lanczos/test_tridiag_forward.py:9-36
    eigvals = np.arange(1.0, 2.0, step=1 / ndim)
    matrix = symmetric_matrix_from_eigenvalues(eigvals).astype(np.float32)
    vector = np.flip(np.arange(1.0, 1.0 + ndim)).astype(np.float32).copy()
    algorithm = lanczos.tridiag(bl.operators.DenseOperator(ndim), ndim, reortho=reortho)
    (vecs, tri), _ = algorithm(vector, matrix)
    Q = vecs.numpy()
    tols = {"atol": 1e-5, "rtol": 1e-5} if reortho == "full" else {"atol": 1e-1, "rtol": 1e-1}
    assert np.allclose(Q.T @ dense_tridiag(*tri) @ Q, matrix, **tols)
    assert np.allclose(Q @ Q.T, np.eye(ndim), **tols)
    assert np.allclose(Q.T @ Q, np.eye(ndim), **tols)


@pytest.mark.parametrize("krylov_depth", [1, 5, 11])
@pytest.mark.parametrize("reortho", ["full", "none"])
def test_mid_rank_reconstruction_satisfies_decomposition(krylov_depth, reortho, ndim=12):
    # /root/reference/tests/test_lanczos/test_tridiag_forward.py:41-58
    eigvals = np.arange(1.0, 2.0, step=1 / ndim)
    matrix = symmetric_matrix_from_eigenvalues(eigvals).astype(np.float32)
    vector = np.flip(np.arange(1.0, 1.0 + ndim)).astype(np.float32).copy()
    algorithm = lanczos.tridiag(bl.operators.DenseOperator(ndim), krylov_depth, reortho=reortho)
    (vecs, tri), (q, b) = algorithm(vector, matrix)
    Q, T = vecs.numpy(), dense_tridiag(*tri)
    e_K = np.eye(krylov_depth)[-1]
    assert np.allclose(matrix @ Q.T, Q.T @ T + np.outer(e_K, q.numpy() * b).T, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("krylov_depth", [1, 5, 10])
@pytest.mark.parametrize("reortho", ["none", "full"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_decomposition_is_satisfied(krylov_depth, reortho, dtype, nrows=10):
    # /root/reference/tests/test_arnoldi/test_hessenberg_forward.py:10-37 (real dtypes)
    rng = np.random.default_rng(1)
    A = rng.standard_normal((nrows, nrows)).astype(dtype)
    v = rng.standard_normal(nrows).astype(dtype)
    Q, H, r, c = arnoldi.hessenberg(bl.operators.DenseOperator(nrows), krylov_depth, reortho=reortho)(v, A)
    assert Q.shape == (nrows, krylov_depth) and H.shape == (krylov_depth, krylov_depth)
    assert r.shape == (nrows,) and c.shape == ()
    Q, H, r, c = Q.numpy(), H.numpy(), r.numpy(), float(c)
    small = np.sqrt(np.finfo(dtype).eps)
    e0, ek = np.eye(krylov_depth)[[0, -1], :]
    assert np.allclose(A @ Q - Q @ H - np.outer(r, ek), 0.0, atol=small, rtol=small)
    assert np.allclose(Q.T @ Q - np.eye(krylov_depth), 0.0, atol=small, rtol=small)
    assert np.allclose(Q @ e0, c * v, atol=small, rtol=small)


@pytest.mark.parametrize("krylov_depth", [1, 5, 10])
def test_reorthogonalisation_improves_the_estimate(krylov_depth, nrows=10):
    # /root/reference/tests/test_arnoldi/test_hessenberg_forward.py:40-66 (Hilbert matrix)
    a = np.arange(nrows)
    A = (1 / (1 + a[:, None] + a[None, :])).astype(np.float32)
    v = np.random.default_rng(2).standard_normal(nrows).astype(np.float32)
    Q, H, r, c = arnoldi.hessenberg(bl.operators.DenseOperator(nrows), krylov_depth, reortho="full")(v, A)
    Q, H, r, c = Q.numpy(), H.numpy(), r.numpy(), float(c)
    small = np.sqrt(np.finfo(np.float32).eps)
    e0, ek = np.eye(krylov_depth)[[0, -1], :]
    assert np.allclose(A @ Q - Q @ H - np.outer(r, ek), 0.0, atol=small, rtol=small)
    assert np.allclose(Q.T @ Q - np.eye(krylov_depth), 0.0, atol=small, rtol=small)
    assert np.allclose(Q @ e0, c * v, atol=small, rtol=small)


@pytest.mark.parametrize("reortho", ["full", "none"])
def test_adjoint_matches_finite_differences(reortho, nrows=15, krylov_depth=10):
    """/root/reference/tests/test_arnoldi/test_hessenberg_adjoint.py:53-99 compares the adjoint with
    autodiff in fp64 (replayed against the stored autodiff VJPs in tests/golden).  Autodiff through
    the loop does not exist here, so this is the independent check: a central finite difference of
    <cotangent, outputs> along random directions, dense cotangents on all of (Q, H, r, c), on a
    well-conditioned operand (on the reference's Hilbert matrix the difference quotient itself is
    only good to a few percent)."""
    rng = np.random.default_rng(2)
    spd = symmetric_matrix_from_eigenvalues(1.0 + rng.uniform(size=nrows), seed=3)
    A = np.tril(spd) - 0.5 * np.diag(np.diag(spd))
    v = rng.standard_normal(nrows)
    alg = arnoldi.hessenberg(bl.operators.DenseOperator(nrows, sym=True), krylov_depth, reortho=reortho)
    (Q, H, r, c), pull = bl.vjp(alg, v, A)
    cot = (rng.standard_normal((nrows, krylov_depth)), rng.standard_normal((krylov_depth, krylov_depth)),
           rng.standard_normal(nrows), rng.standard_normal())  # fmt: skip
    dv, dA = pull(cot)

    def phi(vv, AA):
        Q, H, r, c = alg(vv, AA)
        return (np.sum(cot[0] * Q.numpy()) + np.sum(cot[1] * H.numpy()) + np.sum(cot[2] * r.numpy())
                + cot[3] * float(c))  # fmt: skip

    for _ in range(3):
        dvv, dAA = rng.standard_normal(nrows), rng.standard_normal((nrows, nrows))
        eps = 1e-6
        fd = (phi(v + eps * dvv, A + eps * dAA) - phi(v - eps * dvv, A - eps * dAA)) / (2 * eps)
        an = np.dot(dv.numpy(), dvv) + np.sum(dA.numpy() * dAA)
        if reortho == "full":
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (fd, an)
        else:  # without the re-projection the adjoint is only as exact as the basis is orthogonal
            assert abs(fd - an) <= 1e-4 * max(1.0, abs(an)), (fd, an)


def test_integrand_spd_value_matches_dense_logdet_quadform(n=10):
    """/root/reference/tests/test_lanczos/test_integrand_spd_value_and_grad.py: at full Krylov depth
    the quadrature is exact: v^T log(A) v, and the parameter gradient is that of the dense formula."""
    eigvals = np.arange(0.0, 1.0 + n) + 1.0
    A = symmetric_matrix_from_eigenvalues(eigvals)
    P = np.triu(A) - np.diag(0.5 * np.diag(A))  # _sym(): matvec = (p + p.T) @ x
    rng = np.random.default_rng(2)
    v = rng.integers(0, 2, n + 1) * 2.0 - 1.0
    integrand = lanczos.integrand_spd(np.log, n + 1, bl.operators.DenseOperator(n + 1, sym=True))
    w, U = np.linalg.eigh(A)
    logA = (U * np.log(w)) @ U.T
    assert np.allclose(integrand(v, P), v @ logA @ v, rtol=1e-8)
    value, (_, grad) = integrand.value_and_grad(v, P)
    eps = 1e-6
    dP = np.triu(rng.standard_normal((n + 1, n + 1)))
    fd = (integrand(v, P + eps * dP) - integrand(v, P - eps * dP)) / (2 * eps)
    assert abs(fd - np.sum(grad.numpy() * dP)) < 1e-5 * max(1.0, abs(fd))


def test_custom_vjp_is_similar_but_different(n=3):
    # /root/reference/tests/test_hutchinson.py:9-37 (10 000 probes: forward identical, backward
    # sampled with a different key: different, but within 25 %)
    eigvals = np.arange(0.0, 1.0 + n) + 1.0
    A = symmetric_matrix_from_eigenvalues(eigvals)
    op = bl.operators.DenseOperator(n + 1)
    sampler = hutchinson.sampler_rademacher(np.ones(n + 1), num=2000)
    integrand = lanczos.integrand_spd(np.log, n // 2 + 1, op)
    est_ref = hutchinson.hutchinson(integrand, sampler)
    est_custom = hutchinson.hutchinson_custom_vjp(integrand, sampler)
    key = hutchinson.prng_key(2)
    value_custom, vjp_custom = bl.vjp(est_custom, key, A)
    value_ref, (grad_ref,) = est_ref.value_and_grad(key, A)
    assert np.allclose(value_custom, value_ref)
    _, grad_custom = vjp_custom(1.0)
    g_c, g_r = np.asarray(grad_custom), np.asarray(grad_ref)
    assert not np.allclose(g_c, g_r)
    assert np.linalg.norm(g_c - g_r) < 0.25 * np.linalg.norm(g_r)
    with pytest.raises(RuntimeError, match="oops"):
        est_custom(key, A)


def test_hutchinson_batch_averages_split_keys():
    # /root/reference/src/matfree_extensions/hutchinson.py:57-65
    op = bl.operators.DenseOperator(4)
    A = symmetric_matrix_from_eigenvalues([1.0, 2.0, 3.0, 4.0])
    integrand = lanczos.integrand_spd(np.log, 4, op)
    est = hutchinson.hutchinson_nograd(integrand, hutchinson.sampler_rademacher(np.ones(4), num=50))
    batched = hutchinson.hutchinson_batch(est, num=4)
    key = hutchinson.prng_key(0)
    expected = np.mean([est(k, A) for k in hutchinson.split(key, 4)])
    assert np.allclose(batched(key, A), expected)
    assert abs(batched(key, A) - np.log([1.0, 2.0, 3.0, 4.0]).sum()) < 0.5


# ---- round-2 fixtures (oracle/make_golden_r2.py: the reference's own sources) ------------------------------------
@pytest.mark.parametrize("name", ["suitesparse_1138_bus_k20_f64", "suitesparse_1138_bus_k20_f32"])
def test_suitesparse_file_through_the_gpu_path(name):
    """`exp_util.suite_sparse_load("1138_bus")` (exp_util.py:35-42) -> the benchmark's BCOO operand
    (suite_sparse/benchmark.py:61-68) -> `tridiag(reortho="full")` and its VJP: here
    `SparseOperator.from_matrix_market` on the same file -> the CUDA path.  SPD with cond ~ 8.6e6 (the only
    ill-conditioned sparse fixture the reference ships); parameter gradient in the file's COO order."""
    import os

    from conftest import GOLDEN_DIR, golden, rel_err

    g = golden(name)
    x64 = bool(g["x64"])
    dtype = np.float64 if x64 else np.float32
    K, n = int(g["K"]), int(g["n"])
    op, data = bl.operators.SparseOperator.from_matrix_market(os.path.join(GOLDEN_DIR, "1138_bus.mtx"))
    assert np.array_equal(data, g["data"])
    alg = bl.lanczos.tridiag(op, K, reortho="full")
    ((Qt, (alpha, beta)), (q_rem, b_rem)), pull = bl.vjp(alg, g["v"].astype(dtype), data.astype(dtype))
    t_val, t_grad = (1e-10, 1e-10) if x64 else (1e-5, 1e-4)
    assert rel_err(alpha, g["alpha"]) < t_val and rel_err(beta, g["beta"]) < t_val
    assert rel_err(Qt.numpy(), g["Qt"]) < 10 * t_val
    assert rel_err(q_rem.numpy(), g["q_rem"]) < 10 * t_val and abs(float(b_rem) - float(g["b_rem"])) < 10 * t_val * float(g["b_rem"])
    dv, dp = pull(((g["dQt"].astype(dtype), (g["dalpha"], g["dbeta"])), (g["dq_rem"].astype(dtype), float(g["db_rem"]))))
    assert rel_err(dv.numpy(), g["dv"]) < 5 * t_grad and rel_err(dp.numpy(), g["dp"]) < 5 * t_grad
    dv0, dp0 = pull(((None, (g["dalpha"], g["dbeta"])), (None, None)))  # SLQ-style cotangent
    assert rel_err(dv0.numpy(), g["dv_slqcot"]) < 5 * t_grad and rel_err(dp0.numpy(), g["dp_slqcot"]) < 5 * t_grad
    if not x64:  # and against the float64 run of the reference on the same file (different start vector: values only)
        g64 = golden("suitesparse_1138_bus_k20_f64")
        ((_, (a64, b64)), _), _ = bl.vjp(alg, g64["v"].astype(dtype), data.astype(dtype))
        assert rel_err(a64, g64["alpha"]) < 1e-5 and rel_err(b64, g64["beta"]) < 1e-5


def test_batched_initial_conditions_match_the_reference_vmap():
    """`jax.vmap(solve, in_axes=(0, None))(y0s, scale)` of the PDE training loss
    (/root/reference/experiments/applications/partial_differential_equation/train.py:104-110): three initial conditions,
    one parameter field, lockstep Arnoldi runs (`pde.vmap_solver` -> bl_arnoldi_{forward,adjoint}_batch with per-run
    dQ, dH, dc).  Outputs and dy0 per run, dscale summed over the batch, at the wave fixture's 1e-9."""
    from conftest import golden, rel_err

    from experiments_lanczos_adjoints_b200 import pde

    g = golden("pde_wave_batch_g8_k6_f64")
    gg, K, B, t1 = int(g["g"]), int(g["K"]), int(g["B"]), float(g["t1"])
    field, _like = pde.pde_wave_anisotropic(g["scale"], stencil=g["stencil"])
    solve = pde.solver_expm(0.0, t1, field, expm=pde.expm_arnoldi(K))
    batched = pde.vmap_solver(solve)
    y0s = g["y0s"].reshape(B, -1)
    outs, info = batched(y0s, g["scale"])
    assert info == {"num_matvecs": K}
    assert rel_err(outs.numpy(), g["expm_out"].reshape(B, -1)) < 1e-10
    (outs, _), pullback = bl.vjp(batched, y0s, g["scale"])
    dy0s, dscale = pullback(g["u"].reshape(B, -1))
    assert rel_err(outs.numpy(), g["expm_out"].reshape(B, -1)) < 1e-10
    assert rel_err(dy0s.numpy(), g["loss_dy0s"].reshape(B, -1)) < 1e-9
    assert rel_err(dscale.numpy(), g["loss_dscale"]) < 1e-9
    # the batch equals the runs one at a time
    total = 0.0
    for b in range(B):
        (y1, _), pull1 = bl.vjp(solve, g["y0s"][b], g["scale"])
        d1, ds1 = pull1(g["u"][b])
        assert rel_err(outs.numpy()[b], y1.numpy()) < 1e-12 and rel_err(dy0s.numpy()[b], d1.numpy()) < 1e-11
        total = total + ds1.numpy()
    assert rel_err(dscale.numpy(), total) < 1e-11


@pytest.mark.parametrize("krylov_depth", [1, 5, 10])
@pytest.mark.parametrize("reortho", ["none", "full"])
@pytest.mark.parametrize("ctype", [np.complex64, np.complex128])
def test_decomposition_is_satisfied_for_complex_inputs(krylov_depth, reortho, ctype, nrows=10):
    # /root/reference/tests/test_arnoldi/test_hessenberg_forward.py:10-37 with dtype=complex
    from oracle import krylov, operators

    rng = np.random.default_rng(1)
    A = (rng.standard_normal((nrows, nrows)) + 1j * rng.standard_normal((nrows, nrows))).astype(ctype)
    v = (rng.standard_normal(nrows) + 1j * rng.standard_normal(nrows)).astype(ctype)
    algorithm = arnoldi.hessenberg(bl.operators.DenseOperator(nrows), krylov_depth, reortho=reortho)
    Q, H, r, c = algorithm(v, A)
    assert Q.shape == (nrows, krylov_depth) and H.shape == (krylov_depth, krylov_depth)
    assert r.shape == (nrows,) and np.shape(c) == ()
    Qh, rh = Q.numpy(), r.numpy()
    small_value = np.sqrt(np.finfo(H.dtype).eps)
    tols = {"atol": small_value, "rtol": small_value}
    e0, ek = np.eye(krylov_depth)[[0, -1], :]
    assert np.allclose(A @ Qh - Qh @ H - np.outer(rh, ek), 0.0, **tols)
    assert np.allclose(Qh.T.conj() @ Qh - np.eye(krylov_depth), 0.0, **tols)
    assert np.allclose(Qh @ e0, c * v, **tols)
    # and the oracle's complex restatement (arnoldi.py:66,87,92,95 keep their .conj())
    Q_r, H_r, r_r, c_r = krylov.arnoldi_forward(operators.DenseOperator(), krylov_depth, v.astype(np.complex128),
                                                A.astype(np.complex128))  # fmt: skip
    t = 1e-4 if ctype == np.complex64 else 1e-10
    assert np.abs(H - H_r).max() < t * np.abs(H_r).max() and np.abs(Qh - Q_r).max() < 10 * t
    with pytest.raises(NotImplementedError):
        bl.vjp(algorithm, v, A)
