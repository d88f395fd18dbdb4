"""Round-2 additions to `tests/golden/` (same rules as make_golden.py: the reference's UNMODIFIED sources over the
torch-backed `jax` stand-in; TEST INFRASTRUCTURE, run in the build container only):

    python oracle/make_golden_r2.py

* `pde_wave_batch_g8_k6_f64`: `jax.vmap(solve, in_axes=(0, None))(y0s, scale)` over three initial conditions sharing
  one parameter field -- the batched call of the reference's training loss
  (/root/reference/experiments/applications/partial_differential_equation/train.py:104-110) -- value and gradient.
* `suitesparse_1138_bus_k{K}_{f32,f64}`: `exp_util.suite_sparse_load("1138_bus")` (the SuiteSparse fixture the
  reference ships, SPD, cond ~ 8.6e6; the file is copied next to the fixtures) -> BCOO operand of
  suite_sparse/benchmark.py:61-68 -> `lanczos.tridiag(reortho="full")` forward and VJP, dense and SLQ cotangents.
"""

from __future__ import annotations

import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up sys.path for the shim and the reference)
from make_golden import N, T, jax, jnp, lanczos, save, set_x64, torch  # noqa: E402


def pde_batch_case(name, *, g, K, B, x64, seed):
    from matfree_extensions.util import pde_util

    dt = set_x64(x64)
    rng = np.random.default_rng(seed)
    xs_1d = np.linspace(0.0, 1.0, g)
    dx = xs_1d[1] - xs_1d[0]
    stencil = pde_util.stencil_laplacian(T(dx, dt))
    boundary = pde_util.boundary_neumann()
    scale = 0.2 + 0.05 * rng.standard_normal((g, g))
    pde_rhs, _ = pde_util.pde_wave_anisotropic(T(scale, dt), constrain=jnp.square, stencil=stencil, boundary=boundary)

    def vector_field(x, p):
        return pde_rhs(scale=p)(x)

    t1 = 0.05
    solve = pde_util.solver_expm(0.0, t1, vector_field, expm=pde_util.expm_arnoldi(K))
    y0s = rng.standard_normal((B, 2, g, g))
    u = rng.standard_normal((B, 2, g, g))

    def loss(y_inits, p):
        approx, _aux = jax.vmap(solve, in_axes=(0, None))(y_inits, p)  # train.py:109
        return jnp.sum(approx * T(u, dt))

    val, (dy0s, dp) = jax.value_and_grad(loss, argnums=(0, 1))(T(y0s, dt), T(scale, dt))
    approx, _ = jax.vmap(solve, in_axes=(0, None))(T(y0s, dt), T(scale, dt))
    save(name, g=g, K=K, B=B, dx=dx, stencil=N(stencil), scale=scale, y0s=y0s, u=u, t1=t1, x64=x64,
         expm_out=N(approx), loss=N(val), loss_dy0s=N(dy0s), loss_dscale=N(dp))  # fmt: skip


def suitesparse_case(which, *, K, x64, seed):
    import jax.experimental.sparse  # noqa: F401
    from matfree_extensions.util import exp_util

    dt = set_x64(x64)
    src = f"/root/reference/data/matrices/{which}/{which}.mtx"
    dst = os.path.join(mg.OUT, f"{which}.mtx")
    if not os.path.exists(dst):
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
    M = exp_util.suite_sparse_load(which, path="/root/reference/data/matrices/")  # exp_util.py:35-42
    params, indices = M.data, M.indices
    n = M.shape[0]
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.int32)

    def matvec(x, p):  # benchmark.py:64-68
        return jax.experimental.sparse.BCOO((p, idx), shape=M.shape) @ x

    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n)
    data = np.asarray(N(params), dtype=np.float64)
    alg = lanczos.tridiag(matvec, K, reortho="full", custom_vjp=True)
    out, vjp = jax.vjp(alg, T(v, dt), T(data, dt))
    (Qt, (alpha, beta)), (q_rem, b_rem) = out
    cot = dict(dQt=rng.standard_normal((K, n)), dalpha=rng.standard_normal(K), dbeta=rng.standard_normal(K - 1),
               dq_rem=rng.standard_normal(n), db_rem=rng.standard_normal(()))  # fmt: skip
    ct = ((T(cot["dQt"], dt), (T(cot["dalpha"], dt), T(cot["dbeta"], dt))), (T(cot["dq_rem"], dt), T(cot["db_rem"], dt)))
    dv, dp = vjp(ct)
    z = lambda a: torch.zeros_like(a)  # noqa: E731
    dv0, dp0 = vjp(((z(ct[0][0]), ct[0][1]), (z(ct[1][0]), z(ct[1][1]))))
    tag = "f64" if x64 else "f32"
    save(f"suitesparse_{which}_k{K}_{tag}", which=which, n=n, row=np.asarray(indices)[:, 0], col=np.asarray(indices)[:, 1],
         data=data, v=v, K=K, x64=x64, Qt=N(Qt), alpha=N(alpha), beta=N(beta), q_rem=N(q_rem), b_rem=N(b_rem), **cot,
         dv=N(dv), dp=N(dp), dv_slqcot=N(dv0), dp_slqcot=N(dp0))  # fmt: skip


def main():
    pde_batch_case("pde_wave_batch_g8_k6_f64", g=8, K=6, B=3, x64=True, seed=21)
    suitesparse_case("1138_bus", K=20, x64=True, seed=22)
    suitesparse_case("1138_bus", K=20, x64=False, seed=23)
    set_x64(False)


if __name__ == "__main__":
    main()
