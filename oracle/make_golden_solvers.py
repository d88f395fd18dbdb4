"""Golden vectors of the solver half of the GP path: CG / PCG, partial and pivoted
partial Cholesky, the low-rank preconditioner, and the GP log-marginal likelihood with its gradient.

TEST INFRASTRUCTURE, build container only (needs `/root/reference`):  python oracle/make_golden_solvers.py

The reference's UNMODIFIED `cg.py`, `low_rank.py` and `util/gp_util.py` are imported; `import jax` resolves to
the torch-backed stand-in of `oracle/jaxshim/` (`custom_linear_solve` = primal from the solver, derivative by
the implicit function theorem, like JAX).  Probes are explicit (no JAX PRNG)."""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, HERE)

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
import torch  # noqa: E402
from make_golden import N, T, save, set_x64, spd_from_eigs  # noqa: E402
from matfree_extensions import cg, low_rank  # noqa: E402
from matfree_extensions.util import gp_util  # noqa: E402


def cg_case(name, *, eigs, x64, seed):
    """tests/test_cg/test_cg.py:10-31 (fixed and adaptive CG on a dense SPD matrix)."""
    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    A = spd_from_eigs(np.asarray(eigs, dtype=float), rng)
    n = len(A)
    b = np.arange(1.0, n + 1.0)
    At, bt = T(A, dt), T(b, dt)
    out = dict(A=A, b=b, x64=x64)
    for steps in (n, n // 2, 2 * n):
        x, info = cg.cg_fixed_step(steps)(lambda v: At @ v, bt)
        out[f"x_fixed_{steps}"] = N(x)
        out[f"r_fixed_{steps}"] = N(info["residual_abs"])
    out["steps"] = np.asarray([n, n // 2, 2 * n])
    atol, rtol, maxiter, miniter = 1e-5, 1e-5, 100, 2
    x, info = cg.cg_adaptive(atol=atol, rtol=rtol, maxiter=maxiter, miniter=miniter)(lambda v: At @ v, bt)
    out.update(x_adaptive=N(x), r_adaptive=N(info["residual_abs"]), num_steps=N(info["num_steps"]),
               atol=atol, rtol=rtol, maxiter=maxiter, miniter=miniter)  # fmt: skip
    save(name, **out)


def lowrank_case(name, *, n, rank, x64, seed):
    """tests/test_low_rank/test_low_rank.py (dense matrix elements) + PCG with the pivoted preconditioner
    (test_cg.py:51-83)."""
    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    cov = spd_from_eigs(2.0 ** np.arange(-n // 2, n - n // 2, dtype=float), rng)
    ct = T(cov, dt)

    def element(i, j):
        return ct[i, j]

    L_plain, _ = low_rank.cholesky_partial(rank=rank)(element, n)
    L_pivot, info = low_rank.cholesky_partial_pivot(rank=rank)(element, n)
    b = np.arange(1.0, n + 1.0)
    b /= np.linalg.norm(b)
    small = 1e-2
    pre, _ = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=rank))(element, n)
    Pb = pre(T(b, dt), T(small, dt))
    steps = n // 2
    x, pinfo = cg.pcg_fixed_step(steps)(lambda v: ct @ v + small * v, T(b, dt), lambda v: pre(v, T(small, dt)))
    save(name, cov=cov, n=n, rank=rank, x64=x64, L_plain=N(L_plain), L_pivot=N(L_pivot), success=N(info["success"]),
         b=b, small=small, P_b=N(Pb), pcg_steps=steps, x_pcg=N(x), r_pcg=N(pinfo["residual_abs"]))  # fmt: skip


def logml_case(name, *, n, d, K, rank, num_probes, cg_steps, x64, seed, kind="matern32"):
    """`target_logml(model_gp(mean_constant, kernel), likelihood_pdf_p(gram_matvec(), logpdf_krylov_p(pcg, slq),
    preconditioner(cholesky_partial_pivot)))`: value and gradient w.r.t. every parameter
    (experiments/applications/gaussian_process/train/optim_logml_adjoints_adaptive.py:110-165)."""
    from matfree import hutchinson as mf_hutchinson
    from matfree_extensions import lanczos

    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    y = np.sin(X.sum(-1)) + 0.1 * rng.standard_normal(n)
    probes = rng.integers(0, 2, size=(num_probes, n)) * 2.0 - 1.0
    raw_ls = 0.5 + 0.2 * rng.standard_normal(d)
    raw_os, raw_noise, const = 0.3, -1.0, 0.1
    noise_min = 1e-4

    solve_p = cg.pcg_fixed_step(cg_steps)

    def logdet(A, /, key):  # gp_util.krylov_logdet_slq (num_batches == 1) with explicit probes
        # `A` closes over the kernel parameters.  JAX's closure_convert hoists them into arguments of the
        # custom VJP; the stand-in cannot, so the gradient is taken by autodiff through the Lanczos loop
        # (use_adjoints_for_tridiag=False) -- the same number, as the tridiag/arnoldi fixtures pin
        # ("adjoint == autodiff", tests/test_lanczos/test_tridiag_adjoint.py).
        integrand = lanczos.integrand_spd(jnp.log, K, A, use_adjoints_for_tridiag=False)
        estimate = mf_hutchinson.hutchinson(integrand, lambda _key: T(probes, dt))
        return estimate(key), {"std": 0.0}

    precondition = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=rank))
    logpdf_p = gp_util.logpdf_krylov_p(solve_p=solve_p, logdet=logdet)
    constrain = gp_util.constraint_greater_than(noise_min)
    likelihood, _ = gp_util.likelihood_pdf_p(gp_util.gram_matvec(), logpdf_p, precondition=precondition,
                                             constrain=constrain)  # fmt: skip
    m, _ = gp_util.mean_constant(shape_out=())
    kernels = {"matern32": gp_util.kernel_scaled_matern_32, "rbf": gp_util.kernel_scaled_rbf}
    k, _ = kernels[kind](shape_in=(d,), shape_out=())
    loss = gp_util.target_logml(gp_util.model_gp(m, k), likelihood)

    def mll(ls, os_, rn, cv):
        val, _info = loss(T(X, dt), T(y, dt), None, params_mean={"constant_value": cv},
                          params_kernel={"raw_lengthscale": ls, "raw_outputscale": os_},
                          params_likelihood={"raw_noise": rn})  # fmt: skip
        return val

    args = (T(raw_ls, dt), T(raw_os, dt), T(raw_noise, dt), T(const, dt))
    val, grads = jax.value_and_grad(mll, argnums=(0, 1, 2, 3))(*args)
    save(name, X=X, y=y, probes=probes, K=K, rank=rank, cg_steps=cg_steps, kind=kind, x64=x64, noise_min=noise_min,
         raw_lengthscale=raw_ls, raw_outputscale=raw_os, raw_noise=raw_noise, constant_value=const,
         value=N(val), d_raw_lengthscale=N(grads[0]), d_raw_outputscale=N(grads[1]), d_raw_noise=N(grads[2]),
         d_constant_value=N(grads[3]))  # fmt: skip


def gram_lowrank_case(name, *, n, d, rank, x64, seed):
    """Pivoted partial Cholesky of a Matern-3/2 Gram matrix through the lazy kernel
    (`likelihood_pdf_p.lazy_kernel`, gp_util.py:257-258) and the preconditioner solve."""
    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    raw_ls, raw_os = 0.5 + 0.2 * rng.standard_normal(d), 0.3
    k, _ = gp_util.kernel_scaled_matern_32(shape_in=(d,), shape_out=())
    kernel = k(raw_lengthscale=T(raw_ls, dt), raw_outputscale=T(raw_os, dt))
    Xt = T(X, dt)

    def lazy_kernel(i, j):
        return kernel(Xt[i], Xt[j])

    L, info = low_rank.cholesky_partial_pivot(rank=rank)(lazy_kernel, n)
    v = rng.standard_normal(n)
    noise = 0.05
    pre, _ = low_rank.preconditioner(low_rank.cholesky_partial_pivot(rank=rank))(lazy_kernel, n)
    Pv = pre(T(v, dt), T(noise, dt))
    save(name, X=X, raw_lengthscale=raw_ls, raw_outputscale=raw_os, rank=rank, x64=x64, L=N(L),
         success=N(info["success"]), v=v, noise=noise, P_v=N(Pv))  # fmt: skip


def main():
    cg_case("cg_dense_n9_f64", eigs=np.arange(1.0, 10.0), x64=True, seed=21)
    cg_case("cg_dense_n9_f32", eigs=np.arange(1.0, 10.0), x64=False, seed=22)
    # well conditioned on purpose: an unconverged CG on an ill-conditioned matrix amplifies the summation
    # order of the matvec (1e-4 after 20 steps at cond 3e3), which no restatement can reproduce
    cg_case("cg_dense_n40_f64", eigs=1.0 + 0.5 * np.arange(40.0), x64=True, seed=23)
    lowrank_case("lowrank_dense_n12_r6_f64", n=12, rank=6, x64=True, seed=24)
    lowrank_case("lowrank_dense_n10_r10_f64", n=10, rank=10, x64=True, seed=25)
    lowrank_case("lowrank_dense_n12_r6_f32", n=12, rank=6, x64=False, seed=26)
    gram_lowrank_case("lowrank_gram_n40_d3_r8_f64", n=40, d=3, rank=8, x64=True, seed=27)
    gram_lowrank_case("lowrank_gram_n40_d3_r8_f32", n=40, d=3, rank=8, x64=False, seed=28)
    logml_case("logml_matern32_n30_d3_f64", n=30, d=3, K=8, rank=5, num_probes=6, cg_steps=40, x64=True, seed=29)
    logml_case("logml_rbf_n24_d2_f64", n=24, d=2, K=6, rank=4, num_probes=4, cg_steps=40, x64=True, seed=30, kind="rbf")
    set_x64(False)


if __name__ == "__main__":
    main()
