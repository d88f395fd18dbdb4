"""GPU parity tests of the tcgen05 (tensor-core) Gram kernel: the contraction itself, and the matvec /
VJP against the NumPy oracle evaluated in float64 and against the FP32-ALU kernel, on ragged shapes
(n not a multiple of the 128 x 256 tile, d from 1 to 20).  Tolerances: BASELINE.json north_star fp32
(1e-5 on values, 1e-4 on gradients); the contraction is checked at the fp32 rounding level of its
inputs."""

import numpy as np
import pytest
from conftest import golden, golden_names, rel_err

import experiments_lanczos_adjoints_b200 as bl
from oracle import operators

pytestmark = pytest.mark.gpu


def scaled_f32(X, raw_ls, kind):
    """The scaled inputs as the device holds them (fp32), returned in float64."""
    fac = np.sqrt(3.0) if kind == "matern32" else 1.0  # gp_util.py:84 (Matern-3/2 only)
    return (fac * X / operators.softplus(raw_ls)).astype(np.float32).astype(np.float64)


def problem(n, d, seed, ls_shift=0.5):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    raw_ls = ls_shift + 0.3 * rng.standard_normal(d)
    return X, raw_ls, rng.standard_normal(()), np.asarray(0.05), rng.standard_normal(n), rng.standard_normal(n)


@pytest.mark.parametrize("n,d", [(300, 9), (129, 1), (1000, 20), (2049, 4)])
def test_tensor_core_contraction_matches_float64(n, d):
    """x_i.x_j - |x_j|^2/2 from the TF32 hi/lo split MMA == the float64 value of the same fp32 inputs
    to a few ulps of the largest term (an fp32 dot product has the same bound)."""
    X, raw_ls, raw_os, noise, _, _ = problem(n, d, 0)
    op = bl.operators.GramOperator(X, kind="rbf", path="tensor")
    op.bind((raw_ls, raw_os, noise), np.float32)
    Xs = scaled_f32(X, raw_ls, "rbf")
    for bi, bj in [(0, 0), ((n - 1) // 128, 0)]:
        acc = op.tile_distances(bi, bj).astype(np.float64)
        rows = slice(128 * bi, min(n, 128 * bi + 128))
        cols = slice(256 * bj, min(n, 256 * bj + 256))
        xi, xj = Xs[rows], Xs[cols]
        ref = xi @ xj.T - 0.5 * (xj**2).sum(-1)[None, :]
        scale = np.abs(xi).sum(-1).max() * np.abs(xj).max() + 0.5 * (xj**2).sum(-1).max()
        assert np.abs(acc[: ref.shape[0], : ref.shape[1]] - ref).max() < 8 * 2.0**-24 * scale


@pytest.mark.parametrize("kind", ["matern32", "matern12", "rbf"])
@pytest.mark.parametrize("n,d", [(1, 3), (127, 9), (300, 9), (777, 2), (2500, 16), (4100, 20)])
def test_tensor_core_matvec_and_vjp_match_oracle(kind, n, d):
    X, raw_ls, raw_os, noise, v, lam = problem(n, d, 1)
    orc = operators.GramOperator(X, kind=kind)
    # Matern-1/2 is not differentiable at s = 0: sqrt(s2 + eps) makes the fp32 diagonal differ from
    # the float64 one by sqrt(eps_f32) = 3e-4 (same allowance as tests/test_oracle_golden.py)
    amp = 100.0 if kind == "matern12" else 1.0
    y64 = orc.matvec(v, raw_ls, raw_os, noise)
    z64, (dls64, dos64, dn64) = orc.vjp(v, lam, raw_ls, raw_os, noise)
    res = {}
    for path in ("tensor", "alu"):
        op = bl.operators.GramOperator(X, kind=kind, path=path)
        op.bind((raw_ls, raw_os, noise), np.float32)
        y = op.matvec(bl.asarray(v.astype(np.float32))).numpy()
        op.grad_zero(np.float32)
        z = op.vjp(bl.asarray(v.astype(np.float32)), bl.asarray(lam.astype(np.float32))).numpy()
        dls, dos, dn = (a.numpy() for a in op.grad_export(np.float32))
        res[path] = (y, z, dls, dos)
        assert rel_err(y, y64) < amp * 1e-5
        assert rel_err(z, z64) < amp * 1e-5
        if kind != "matern12" and n > 1:
            assert rel_err(dls, dls64) < 1e-4
        assert rel_err(dos, dos64) < amp * 1e-4
        assert rel_err(dn, dn64) < 1e-4
    # the two kernels agree with each other to fp32 rounding of the distances
    assert rel_err(res["tensor"][0], res["alu"][0]) < amp * 1e-5
    assert rel_err(res["tensor"][1], res["alu"][1]) < amp * 1e-5


def test_tensor_core_diagonal_is_exact_for_duplicate_free_inputs():
    """k(x_i, x_i): s2 = 0 exactly on the diagonal (as the fp32 evaluation of the reference's expanded
    form gives), so with a tiny lengthscale the Gram matrix is sigma (1 + sqrt(eps)) e^{-sqrt(eps)} I."""
    n, d = 515, 9
    rng = np.random.default_rng(2)
    X = 100.0 * rng.standard_normal((n, d))  # scaled distances >> 1: off-diagonal entries underflow
    op = bl.operators.GramOperator(X, kind="matern32", path="tensor")
    raw_os = np.asarray(0.3)
    v = rng.standard_normal(n).astype(np.float32)
    y = op(v, np.zeros(d), raw_os, np.asarray(0.0)).numpy()
    s = np.sqrt(np.float32(np.finfo(np.float32).eps))
    expect = operators.softplus(raw_os) * (1 + s) * np.exp(-s) * v
    assert rel_err(y, expect) < 1e-6


@pytest.mark.parametrize("name", golden_names("gp_kernels_"))
def test_tensor_core_path_is_the_default_for_fp32(name):
    """`path="auto"` == `path="tensor"` bit for bit in fp32 (the golden test above therefore pins the
    tensor-core kernel); fp64 always takes the FP64-ALU kernel."""
    g = golden(name)
    X, v = g["X"], g["v"].astype(np.float32)
    params = (g["raw_lengthscale"].astype(np.float32), g["raw_outputscale"].astype(np.float32), np.zeros((), np.float32))
    y_auto = bl.operators.GramOperator(X, kind="matern32")(v, *params).numpy()
    y_tc = bl.operators.GramOperator(X, kind="matern32", path="tensor")(v, *params).numpy()
    assert np.array_equal(y_auto, y_tc)


def test_tensor_core_path_rejects_large_dimension():
    X = np.random.default_rng(0).standard_normal((64, 24))
    with pytest.raises(ValueError):  # BL_EINVAL
        bl.operators.GramOperator(X, path="tensor")
    op = bl.operators.GramOperator(X)  # automatic: falls back to the ALU kernel for d > 20
    y = op(np.ones(64, np.float32), np.zeros(24), np.asarray(0.0), np.asarray(0.1)).numpy()
    ref = operators.GramOperator(X).matvec(np.ones(64), np.zeros(24), np.asarray(0.0), np.asarray(0.1))
    assert rel_err(y, ref) < 1e-5


@pytest.mark.parametrize("kind", ["matern32", "rbf", "matern12"])
@pytest.mark.parametrize("n,d,K", [(700, 9, 10), (3000, 4, 20), (257, 16, 7)])
def test_deferred_batched_parameter_cotangent_matches_per_step_sweeps(kind, n, d, K, monkeypatch):
    """The adjoint sweep defers the Gram operator's parameter cotangent to ONE batched pass over the K
    (lambda_idx, q_idx) pairs (`vjp_batch`, more than 16 pairs -> several passes); same gradient as K
    per-step cotangent sweeps (BL_GRAM_DEFER=0) and as the float64 oracle."""
    from oracle import krylov

    X, raw_ls, raw_os, noise, v, _ = problem(n, d, 5, ls_shift=1.0)
    rng = np.random.default_rng(6)
    dalpha, dbeta = rng.standard_normal(K), rng.standard_normal(K - 1)
    params = (raw_ls.astype(np.float32), np.float32(raw_os), np.float32(0.3))
    out = {}
    for defer in ("1", "0"):
        monkeypatch.setenv("BL_GRAM_DEFER", defer)
        op = bl.operators.GramOperator(X, kind=kind)
        alg = bl.lanczos.tridiag(op, K, reortho="full")
        _, pull = bl.vjp(alg, v.astype(np.float32), *params)
        grads = pull(((None, (dalpha, dbeta)), (None, None)))
        out[defer] = [np.asarray(g.numpy(), np.float64).ravel() for g in grads[1:]]
    for a, b in zip(out["1"], out["0"]):
        assert rel_err(a, b) < 2e-4
    ref = krylov.tridiag(operators.GramOperator(X, kind=kind), K, reortho="full")
    ((Qt, (a_r, b_r)), (q_r, br_r)), pull_r = ref.vjp(v, raw_ls, raw_os, 0.3)
    z = np.zeros_like
    g_r = pull_r(((z(Qt), (dalpha, dbeta)), (z(q_r), z(br_r))))[1:]
    amp = 50.0 if kind == "matern12" else 1.0
    for a, b in zip(out["1"], g_r):
        assert rel_err(a, np.asarray(b).ravel()) < amp * 1e-4


@pytest.mark.parametrize("kind,n,d,K,P", [("matern32", 900, 9, 8, 5), ("rbf", 300, 3, 6, 19), ("matern32", 2100, 4, 10, 16)])
def test_lockstep_probe_batch_matches_sequential_probes(kind, n, d, K, P, monkeypatch):
    """Hutchinson over P probes with the Lanczos runs in lockstep (`bl_arnoldi_*_batch`: one batched Gram
    sweep per step for all probes, one batched cotangent pass) == P sequential runs == the float64 oracle."""
    from oracle import krylov

    X, raw_ls, raw_os, noise, _, _ = problem(n, d, 7, ls_shift=1.0)
    rng = np.random.default_rng(8)
    probes = (rng.integers(0, 2, size=(P, n)) * 2 - 1).astype(np.float32)
    params = (raw_ls.astype(np.float32), np.float32(raw_os), np.float32(0.3))
    out = {}
    for mode in ("batch", "sequential"):
        monkeypatch.setenv("BL_GRAM_DEFER", "1" if mode == "batch" else "0")
        op = bl.operators.GramOperator(X, kind=kind)
        est = bl.hutchinson.hutchinson(bl.lanczos.integrand_spd(np.log, K, op), lambda key: probes)
        launches = bl.launch_count()
        value, grads = est.value_and_grad(None, *params)
        out[mode] = (float(value), [np.asarray(g.numpy(), np.float64).ravel() for g in grads], bl.launch_count() - launches)
        assert abs(float(est(None, *params)) - float(value)) < 1e-5 * abs(float(value))
    assert out["batch"][2] < out["sequential"][2]  # fewer launches: the probes share their Gram sweeps
    assert abs(out["batch"][0] - out["sequential"][0]) < 1e-5 * abs(out["sequential"][0])
    for a, b in zip(out["batch"][1], out["sequential"][1]):
        assert rel_err(a, b) < 2e-4
    integrand = krylov.IntegrandSPD(np.log, lambda x: 1.0 / x, K, operators.GramOperator(X, kind=kind))
    v_r, g_r = krylov.hutchinson_value_and_grad(integrand, probes.astype(np.float64), raw_ls, raw_os, 0.3)
    assert abs(out["batch"][0] - v_r) < 1e-5 * abs(v_r)
    for a, b in zip(out["batch"][1], g_r):
        assert rel_err(a, np.asarray(b).ravel()) < 1e-4
