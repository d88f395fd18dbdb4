"""CPU oracle, solver half of the GP path: CG / PCG, partial (pivoted) Cholesky, the low-rank
preconditioner and the GP log-marginal likelihood with its gradient.

TEST INFRASTRUCTURE ONLY (see `oracle/__init__.py`).  NumPy restatement of
`/root/reference/src/matfree_extensions/{cg,low_rank}.py` and of the model plumbing in
`util/gp_util.py`; every function cites the lines it follows.  Pinned against
`tests/golden/{cg,lowrank,logml}_*.npz`, produced by the reference's own sources
(`oracle/make_golden_solvers.py`).
"""

from __future__ import annotations

import numpy as np

from oracle import krylov
from oracle.operators import GramOperator, softplus, softplus_grad


def safe_divide(a, b):
    """`cg._safe_divide` (`cg.py:196-213`): `a / b` where `|b| > eps^2`, else `a`."""
    eps = np.finfo(np.asarray(a).dtype).eps ** 2
    return a / b if abs(b) > eps else a


def pcg_fixed_step(A, b, P, num_matvecs):
    """`cg.pcg_fixed_step(num_matvecs)(A, b, P)` (`cg.py:20-62`).  Returns `(x, residual)`."""
    b = np.asarray(b)
    x = np.zeros_like(b)
    r = b - A(x)
    z = P(r)
    p = z
    for _ in range(num_matvecs):
        x, p, r, z = _pcg_body(A, P, x, p, r, z)
    return x, r


def _pcg_body(A, P, x, p, r, z):
    """One iteration (`cg.py:42-58`)."""
    Ap = A(p)
    a = safe_divide(np.dot(r, z), np.dot(p, Ap))
    x = x + a * p
    rold, zold = r, z
    r = r - a * Ap
    z = P(r)
    bb = safe_divide(np.dot(r, z), np.dot(rold, zold))
    p = z + bb * p
    return x, p, r, z


def pcg_adaptive(A, b, P, *, atol, rtol, maxiter, miniter):
    """`cg.pcg_adaptive` (`cg.py:75-131`): iterate while `rms(r / (atol + |x| rtol)) > 1` or fewer than
    `miniter` steps were taken, and at most `maxiter` steps.  Returns `(x, residual, num_steps)`."""
    b = np.asarray(b)
    x = np.zeros_like(b)
    r = b - A(x)
    z = P(r)
    p = z
    nsteps = 0
    while True:
        error_rel = r / (atol + np.abs(x) * rtol)
        proceed = (np.sqrt(np.mean(error_rel**2)) > 1.0) or (nsteps < miniter)
        if not (proceed and nsteps < maxiter):
            break
        x, p, r, z = _pcg_body(A, P, x, p, r, z)
        nsteps += 1
    return x, r, nsteps


def cholesky_partial(element_column, diagonal, n, rank):
    """`low_rank.cholesky_partial` (`low_rank.py:63-118`).  `element_column(i)` is column `i` of the
    matrix, `diagonal()` its diagonal."""
    diag = diagonal()
    L = np.zeros((n, rank), dtype=diag.dtype)
    for i in range(rank):
        l_ii = np.sqrt(diag[i] - np.dot(L[i], L[i]))
        L[:, i] = (element_column(i) - L @ L[i, :]) / l_ii
    return L


def cholesky_partial_pivot(element_column, diagonal, n, rank):
    """`low_rank.cholesky_partial_pivot` (`low_rank.py:120-225`) without physically permuting: `perm` is
    the reference's `P_matrix`; the arg-max runs over POSITIONS of the permuted arrangement (first
    maximum, like `jnp.argmax`), the factor is returned in the original row order (`_pivot_invert`).
    Returns `(L, success)`."""
    diag = diagonal()
    L = np.zeros((n, rank), dtype=diag.dtype)  # rows in ORIGINAL order
    perm = np.arange(n)
    success = True
    for i in range(rank):
        res = np.abs(diag[perm] - np.einsum("jk,jk->j", L[perm], L[perm]))  # :177-178
        k = int(np.argmax(res))  # :179
        perm[[i, k]] = perm[[k, i]]  # :182-184
        piv = perm[i]
        l_ii_sq = diag[piv] - np.dot(L[piv], L[piv])  # :194
        with np.errstate(invalid="ignore"):
            l_ii = np.sqrt(l_ii_sq)
        L[:, i] = (element_column(piv) - L @ L[piv, :]) / l_ii  # :196-197 (all rows; same values as permuted)
        success = success and bool(l_ii_sq > 0.0)  # :198
    return L, success


def preconditioner_solve(L, v, s):
    """`low_rank.preconditioner(...).solve(v, s)` (`low_rank.py:36-47`): `(s I + L L^T)^{-1} v` by the
    Woodbury identity with a Cholesky of the capacitance matrix."""
    U = L / np.sqrt(s)
    V = L.T / np.sqrt(s)
    v = v / s
    cap = np.eye(L.shape[1], dtype=L.dtype) + V @ U
    c = np.linalg.cholesky(cap)
    sol = np.linalg.solve(c.T, np.linalg.solve(c, V @ v))
    return v - U @ sol


class GPLogML:
    """`target_logml(model_gp(mean_constant, kernel), likelihood_pdf_p(gram_matvec(), logpdf_krylov_p(pcg,
    slq), preconditioner(cholesky_partial_pivot)))` (`gp_util.py:15-33, 243-276, 414-431`, wiring of
    `optim_logml_adjoints_adaptive.py:110-140`) for explicit probes.

    Parameters: `raw_lengthscale (d,)`, `raw_outputscale`, `raw_noise`, `constant_value`; the noise is
    `noise_min + softplus(raw_noise)` (`gp_util.constraint_greater_than`, `gp_util.py:187-201`)."""

    def __init__(self, X, y, *, kind, krylov_depth, probes, rank, cg_steps, noise_min):
        self.X, self.y, self.kind = np.asarray(X), np.asarray(y), kind
        self.K, self.probes, self.rank, self.cg_steps = krylov_depth, np.asarray(probes), rank, cg_steps
        self.noise_min = noise_min
        self.op = GramOperator(self.X, kind=kind)

    def _kernel_matrix(self, raw_ls, raw_os):
        n = len(self.X)
        return np.stack([self.op.matvec(e, raw_ls, raw_os, 0.0) for e in np.eye(n, dtype=self.X.dtype)], axis=1)

    def value_and_grad(self, raw_ls, raw_os, raw_noise, const):
        n = len(self.y)
        noise = self.noise_min + softplus(raw_noise)
        Kmat = self._kernel_matrix(raw_ls, raw_os)  # lazy_kernel WITHOUT noise (gp_util.py:257-258)
        L, _ = cholesky_partial_pivot(lambda i: Kmat[:, i], lambda: np.diag(Kmat).copy(), n, self.rank)

        def A(v):
            return self.op.matvec(v, raw_ls, raw_os, noise)  # cov_matvec(v) + noise * v   (:270)

        # log-determinant by SLQ (gp_util.krylov_logdet_slq, one batch) and its parameter gradient
        integrand = krylov.IntegrandSPD(np.log, lambda x: 1.0 / x, self.K, _BoundNoise(self.op))
        logdet, (g_ls, g_os, g_noise) = krylov.hutchinson_value_and_grad(integrand, self.probes, raw_ls, raw_os, noise)
        # Mahalanobis term (gp_util.py:421-423); gradient by the implicit function theorem
        # (custom_linear_solve, cg.py:25-27): d/dtheta [r^T A^{-1} r] = -alpha^T dA alpha, alpha = A^{-1} r
        r = self.y - const
        alpha, _ = pcg_fixed_step(A, r, lambda v: preconditioner_solve(L, v, noise), self.cg_steps)
        maha = np.dot(r, alpha)
        _, (m_ls, m_os, m_noise) = self.op.vjp(alpha, alpha, raw_ls, raw_os, noise)
        value = -0.5 * logdet - 0.5 * maha - n / 2 * np.log(2 * np.pi)  # :428
        d_ls = -0.5 * g_ls + 0.5 * m_ls
        d_os = -0.5 * g_os + 0.5 * m_os
        d_noise = (-0.5 * g_noise + 0.5 * m_noise) * softplus_grad(raw_noise)
        # d/dconst [-0.5 (y-c)^T A^{-1} (y-c)] = sum(alpha) (A symmetric); the solve's own residual is ignored
        d_const = np.sum(alpha)
        return value, (d_ls, d_os, d_noise, d_const)


class _BoundNoise:
    """Adapter: the Gram oracle operator with `noise` as an ordinary third parameter."""

    num_params = 3

    def __init__(self, op):
        self.op = op

    def matvec(self, v, raw_ls, raw_os, noise):
        return self.op.matvec(v, raw_ls, raw_os, noise)

    def vjp(self, q, lam, raw_ls, raw_os, noise):
        return self.op.vjp(q, lam, raw_ls, raw_os, noise)

    def __call__(self, v, *params):
        return self.matvec(v, *params)
