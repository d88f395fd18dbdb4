"""Generate `tests/golden/*.npz` by running the reference's UNMODIFIED sources.

TEST INFRASTRUCTURE.  Run in the build container only (needs `/root/reference`):

    python oracle/make_golden.py

`/root/reference/src/matfree_extensions/{arnoldi,lanczos,hutchinson}.py` and the kernels /
stencil in `util/{gp_util,pde_util}.py` are imported as they are; `import jax` resolves to
the torch-backed stand-in in `oracle/jaxshim/` (JAX itself is not installed here — see
`oracle/jaxshim/jax/_core.py`).  Every fixture stores its inputs explicitly (drawn with
NumPy, never with a JAX PRNG), the reference's outputs, and — where the reference's own
tests compare the custom VJP against autodiff — both VJPs.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference/src")

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
import torch  # noqa: E402
from matfree_extensions import arnoldi, hutchinson, lanczos  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def T(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def N(x):
    if isinstance(x, (tuple, list)):
        return [N(e) for e in x]
    return np.asarray(x.detach()) if isinstance(x, torch.Tensor) else np.asarray(x)


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", os.path.relpath(path), sorted(arrays))


def set_x64(flag):
    jax.config.update("jax_enable_x64", flag)
    return torch.float64 if flag else torch.float32


def hilbert(n):
    a = np.arange(n)
    return 1.0 / (1 + a[:, None] + a[None, :])


def lower_half(m):
    t = np.tril(m)
    return t - 0.5 * np.diag(np.diag(t))


def upper_half(m):
    return np.triu(m) - 0.5 * np.diag(np.diag(m))


def spd_from_eigs(eigs, rng):
    U, _ = np.linalg.qr(rng.standard_normal((len(eigs), len(eigs))))
    return (U * eigs) @ U.T


MATVECS = {
    "dense": lambda s, p: p @ s,
    "sym": lambda s, p: (p + p.T) @ s,
}


def arnoldi_case(name, *, A, v, K, reortho, matvec, x64, seed, reortho_vjp="match"):
    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    n = len(v)
    cot = dict(
        dQ=rng.standard_normal((n, K)), dH=rng.standard_normal((K, K)),
        dr=rng.standard_normal(n), dc=rng.standard_normal(()),
    )  # fmt: skip
    mv = MATVECS[matvec]
    kw = {"reortho": reortho, "reortho_vjp": reortho_vjp}
    alg_adj = arnoldi.hessenberg(mv, K, custom_vjp=True, **kw)
    alg_ad = arnoldi.hessenberg(mv, K, custom_vjp=False, **kw)
    vt, At = T(v, dt), T(A, dt)
    (Q, H, r, c), vjp_adj = jax.vjp(alg_adj, vt, At)
    _, vjp_ad = jax.vjp(alg_ad, vt, At)
    ct = tuple(T(cot[k], dt) for k in ("dQ", "dH", "dr", "dc"))
    dv1, dp1 = vjp_adj(ct)
    dv2, dp2 = vjp_ad(ct)
    save(
        name, A=A, v=v, K=K, reortho=reortho, reortho_vjp=reortho_vjp, matvec=matvec, x64=x64,
        Q=N(Q), H=N(H), r=N(r), c=N(c), **cot,
        dv_adjoint=N(dv1), dp_adjoint=N(dp1), dv_autodiff=N(dv2), dp_autodiff=N(dp2),
    )  # fmt: skip


def tridiag_case(name, *, A, v, K, reortho, matvec, x64, seed):
    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    n = len(v)
    mv = MATVECS[matvec]
    alg_adj = lanczos.tridiag(mv, K, reortho=reortho, custom_vjp=True)
    alg_ad = lanczos.tridiag(mv, K, reortho=reortho, custom_vjp=False)
    vt, At = T(v, dt), T(A, dt)
    out, vjp_adj = jax.vjp(alg_adj, vt, At)
    _, vjp_ad = jax.vjp(alg_ad, vt, At)
    (Qt, (alpha, beta)), (q_rem, b_rem) = out
    cot = dict(
        dQt=rng.standard_normal((K, n)), dalpha=rng.standard_normal(K),
        dbeta=rng.standard_normal(K - 1), dq_rem=rng.standard_normal(n),
        db_rem=rng.standard_normal(()),
    )  # fmt: skip
    ct = (
        (T(cot["dQt"], dt), (T(cot["dalpha"], dt), T(cot["dbeta"], dt))),
        (T(cot["dq_rem"], dt), T(cot["db_rem"], dt)),
    )
    dv1, dp1 = vjp_adj(ct)
    dv2, dp2 = vjp_ad(ct)
    save(
        name, A=A, v=v, K=K, reortho=reortho, matvec=matvec, x64=x64,
        Qt=N(Qt), alpha=N(alpha), beta=N(beta), q_rem=N(q_rem), b_rem=N(b_rem), **cot,
        dv_adjoint=N(dv1), dp_adjoint=N(dp1), dv_autodiff=N(dv2), dp_autodiff=N(dp2),
    )  # fmt: skip


def slq_case(name, *, A, probes, K, matvec, x64):
    dt = set_x64(x64)
    mv = MATVECS[matvec]
    At, Pt = T(A, dt), T(probes, dt)
    res = {}
    for tag, use_adj in (("adjoint", True), ("autodiff", False)):
        integrand = lanczos.integrand_spd(jnp.log, K, mv, use_adjoints_for_tridiag=use_adj)
        estimate = hutchinson.hutchinson_nograd(integrand, lambda key: Pt)
        value, grad = jax.value_and_grad(estimate, argnums=1)(None, At)
        res[f"value_{tag}"], res[f"grad_{tag}"] = N(value), N(grad)
        per_probe = [jax.value_and_grad(integrand, argnums=(0, 1))(p, At) for p in Pt]
        res[f"probe_values_{tag}"] = np.stack([N(v) for v, _ in per_probe])
        res[f"probe_dv0_{tag}"] = np.stack([N(g[0]) for _, g in per_probe])
    integrand = lanczos.integrand_spd_custom_vjp_reuse(jnp.log, K, mv)
    estimate = hutchinson.hutchinson_nograd(integrand, lambda key: Pt)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        value, grad = jax.value_and_grad(estimate, argnums=1)(None, At)
    res["value_reuse"], res["grad_reuse"] = N(value), N(grad)
    batched = hutchinson.hutchinson_batch(
        lambda key, p: jnp.sum(key.to(p.dtype)) * jnp.sum(p), num=3
    )
    del batched  # jax.random.split is not threefry here: nothing to pin
    save(name, A=A, probes=probes, K=K, matvec=matvec, x64=x64, **res)


def sparse_case(name, *, n, nnz_off, K, x64, seed):
    """BCOO operand as in suite_sparse/benchmark.py:61-68; params = COO data, with a
    duplicated entry to pin "duplicates are summed, each is its own parameter"."""
    import jax.experimental.sparse

    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    r = rng.integers(0, n, nnz_off)
    c = rng.integers(0, n, nnz_off)
    keep = r > c
    r, c = r[keep], c[keep]
    vals = -rng.uniform(0.1, 1.0, len(r))
    # symmetric file layout: diagonal + strict lower, then mirrored (mmread order)
    row = np.concatenate([np.arange(n), r, c, [3]])
    col = np.concatenate([np.arange(n), c, r, [3]])
    data = np.concatenate([np.full(n, 8.0) + rng.uniform(0, 1, n), vals, vals, [0.25]])
    idx = torch.as_tensor(np.stack([row, col]).T, dtype=torch.int32)

    def matvec(x, p):
        return jax.experimental.sparse.BCOO((p, idx), shape=(n, n)) @ x

    v = rng.standard_normal(n)
    for reortho in ("full", "none"):
        alg = lanczos.tridiag(matvec, K, reortho=reortho, custom_vjp=True)
        out, vjp = jax.vjp(alg, T(v, dt), T(data, dt))
        (Qt, (alpha, beta)), (q_rem, b_rem) = out
        cot = dict(
            dQt=rng.standard_normal((K, n)), dalpha=rng.standard_normal(K),
            dbeta=rng.standard_normal(K - 1), dq_rem=rng.standard_normal(n),
            db_rem=rng.standard_normal(()),
        )  # fmt: skip
        ct = (
            (T(cot["dQt"], dt), (T(cot["dalpha"], dt), T(cot["dbeta"], dt))),
            (T(cot["dq_rem"], dt), T(cot["db_rem"], dt)),
        )
        dv, dp = vjp(ct)
        # SLQ-style sparse cotangent (only alpha/beta)
        z = lambda a: torch.zeros_like(a)  # noqa: E731
        ct0 = ((z(ct[0][0]), ct[0][1]), (z(ct[1][0]), z(ct[1][1])))
        dv0, dp0 = vjp(ct0)
        save(
            f"{name}_{reortho}", n=n, row=row, col=col, data=data, v=v, K=K, reortho=reortho,
            x64=x64, Qt=N(Qt), alpha=N(alpha), beta=N(beta), q_rem=N(q_rem), b_rem=N(b_rem),
            **cot, dv=N(dv), dp=N(dp), dv_slqcot=N(dv0), dp_slqcot=N(dp0),
        )  # fmt: skip


def gp_case(name, *, n, d, x64, seed):
    from matfree_extensions.util import gp_util

    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d))
    v = rng.standard_normal(n)
    lam = rng.standard_normal(n)
    raw_ls = rng.standard_normal(d)
    raw_os = rng.standard_normal(())
    out = dict(X=X, v=v, lam=lam, raw_lengthscale=raw_ls, raw_outputscale=raw_os, x64=x64)
    kinds = {
        "matern32": gp_util.kernel_scaled_matern_32,
        "matern12": gp_util.kernel_scaled_matern_12,
        "rbf": gp_util.kernel_scaled_rbf,
    }
    for kind, make in kinds.items():
        parametrize, _ = make(shape_in=(d,), shape_out=())

        def mv(vec, ls, os_):
            k = parametrize(raw_lengthscale=ls, raw_outputscale=os_)
            return gp_util.gram_matvec()(k)(T(X, dt), T(X, dt), vec)

        y, vjp = jax.vjp(mv, T(v, dt), T(raw_ls, dt), T(raw_os, dt))
        dvec, dls, dos = vjp(T(lam, dt))
        out.update({f"{kind}_y": N(y), f"{kind}_dv": N(dvec), f"{kind}_dls": N(dls), f"{kind}_dos": N(dos)})
    sp = gp_util.constraint_greater_than(0.0)
    xs = np.array([-30.0, -2.0, 0.0, 3.0, 19.9, 20.0, 25.0])
    out["softplus_x"], out["softplus_y"] = xs, N(sp(T(xs, dt)))
    save(name, **out)


def pde_case(name, *, g, K, x64, seed):
    from matfree_extensions.util import pde_util

    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    xs_1d = np.linspace(0.0, 1.0, g)
    dx = xs_1d[1] - xs_1d[0]
    stencil = pde_util.stencil_laplacian(T(dx, dt))
    boundary = pde_util.boundary_neumann()
    scale = 0.2 + 0.05 * rng.standard_normal((g, g))
    pde_rhs, _ = pde_util.pde_wave_anisotropic(
        T(scale, dt), constrain=jnp.square, stencil=stencil, boundary=boundary
    )

    def vector_field(x, p):
        return pde_rhs(scale=p)(x)

    y0 = rng.standard_normal((2, g, g))
    lam = rng.standard_normal((2, g, g))
    # the operator and its VJP
    y, vjp = jax.vjp(vector_field, T(y0, dt), T(scale, dt))
    dx_, dscale = vjp(T(lam, dt))
    # expm action through the Arnoldi adjoint
    t1 = 0.05
    expm = pde_util.expm_arnoldi(K)
    solve = pde_util.solver_expm(0.0, t1, vector_field, expm=expm)
    u = rng.standard_normal((2, g, g))

    def loss(y_init, p):
        approx, _ = solve(y_init, p)
        return jnp.sum(approx * T(u, dt))

    val, (dy0, dp) = jax.value_and_grad(loss, argnums=(0, 1))(T(y0, dt), T(scale, dt))
    approx, _ = solve(T(y0, dt), T(scale, dt))
    save(
        name, g=g, K=K, dx=dx, stencil=N(stencil), scale=scale, y0=y0, lam=lam, x64=x64,
        rhs=N(y), rhs_dx=N(dx_), rhs_dscale=N(dscale), t1=t1, u=u, expm_out=N(approx),
        loss=N(val), loss_dy0=N(dy0), loss_dscale=N(dp),
    )  # fmt: skip


def main():
    rng = np.random.default_rng(20240518)
    # --- arnoldi: /root/reference/tests/test_arnoldi/test_hessenberg_{forward,adjoint}.py
    for reortho in ("none", "full"):
        arnoldi_case(f"arnoldi_dense_n3_k2_{reortho}_f32", A=rng.standard_normal((3, 3)),
                     v=rng.standard_normal(3), K=2, reortho=reortho, matvec="dense", x64=False, seed=3)  # fmt: skip
        for K in (1, 5, 10):
            arnoldi_case(f"arnoldi_dense_n10_k{K}_{reortho}_f64", A=rng.standard_normal((10, 10)),
                         v=rng.standard_normal(10), K=K, reortho=reortho, matvec="dense", x64=True, seed=K)  # fmt: skip
    arnoldi_case("arnoldi_hilbert_n15_k10_full_f64", A=lower_half(hilbert(15)),
                 v=rng.standard_normal(15), K=10, reortho="full", matvec="sym", x64=True, seed=3)  # fmt: skip
    arnoldi_case("arnoldi_hilbert_n10_k5_full_f32", A=hilbert(10),
                 v=rng.standard_normal(10), K=5, reortho="full", matvec="dense", x64=False, seed=4)  # fmt: skip
    arnoldi_case("arnoldi_dense_n10_k5_fwdnone_f64", A=rng.standard_normal((10, 10)),
                 v=rng.standard_normal(10), K=5, reortho="none", reortho_vjp="none",
                 matvec="dense", x64=True, seed=5)  # fmt: skip

    # --- tridiag: /root/reference/tests/test_lanczos/test_tridiag_{forward,adjoint}.py
    eigs = rng.uniform(size=10) + 1.0
    A10 = spd_from_eigs(eigs, rng)
    for reortho in ("full", "none"):
        tridiag_case(f"tridiag_sym_n10_k4_{reortho}_f32", A=upper_half(A10), v=rng.standard_normal(10),
                     K=4, reortho=reortho, matvec="sym", x64=False, seed=4)  # fmt: skip
        tridiag_case(f"tridiag_sym_n10_k4_{reortho}_f64", A=upper_half(A10), v=rng.standard_normal(10),
                     K=4, reortho=reortho, matvec="sym", x64=True, seed=5)  # fmt: skip
    A12 = spd_from_eigs(np.arange(1.0, 2.0, 1 / 12), rng)
    for K in (1, 5, 11, 12):
        for reortho in ("full", "none"):
            if K == 12 and reortho == "none":
                continue  # b_K ~ 0 blows up the 3-term backward (lanczos.py:183-184)
            if K == 1:
                continue  # K-1 = 0 off-diagonals: covered by the arnoldi K=1 cases
            tridiag_case(f"tridiag_dense_n12_k{K}_{reortho}_f64", A=A12, v=np.flip(np.arange(1.0, 13.0)).copy(),
                         K=K, reortho=reortho, matvec="dense", x64=True, seed=K)  # fmt: skip
    # BASELINE config 1: dense SPD 100x100, K=10, fp64
    A100 = spd_from_eigs(1.0 + rng.uniform(size=100), rng)
    for reortho in ("full", "none"):
        tridiag_case(f"tridiag_sym_n100_k10_{reortho}_f64", A=upper_half(A100), v=rng.standard_normal(100),
                     K=10, reortho=reortho, matvec="sym", x64=True, seed=6)  # fmt: skip

    # --- SLQ: /root/reference/tests/test_lanczos/test_integrand_spd_value_and_grad.py
    A11 = spd_from_eigs(np.arange(0.0, 11.0) + 1.0, rng)
    probes = rng.integers(0, 2, size=(8, 11)) * 2.0 - 1.0
    slq_case("slq_sym_n11_k6_f64", A=upper_half(A11), probes=probes, K=6, matvec="sym", x64=True)
    slq_case("slq_dense_n11_k5_f32", A=A11, probes=probes, K=5, matvec="dense", x64=False)

    # --- sparse operand: suite_sparse/benchmark.py:61-68
    sparse_case("sparse_coo_n60_k8_f64", n=60, nnz_off=400, K=8, x64=True, seed=7)
    sparse_case("sparse_coo_n60_k8_f32", n=60, nnz_off=400, K=8, x64=False, seed=8)

    # --- GP kernels and wave stencil (util/gp_util.py, util/pde_util.py)
    gp_case("gp_kernels_n40_d3_f64", n=40, d=3, x64=True, seed=9)
    gp_case("gp_kernels_n40_d3_f32", n=40, d=3, x64=False, seed=10)
    pde_case("pde_wave_g8_k6_f64", g=8, K=6, x64=True, seed=11)
    set_x64(False)


if __name__ == "__main__":
    main()
