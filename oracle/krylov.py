"""Oracle Krylov core (NumPy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates, in NumPy and with explicit Python loops, the arithmetic of

* `/root/reference/src/matfree_extensions/arnoldi.py`   (Arnoldi + adjoint),
* `/root/reference/src/matfree_extensions/lanczos.py`   (tridiag, 3-term Lanczos + adjoint,
  SLQ integrands),
* `/root/reference/src/matfree_extensions/hutchinson.py` (Hutchinson mean).

Shapes follow the reference (`Q` is `(n, K)`, `tridiag` returns `Q.T`).  An operator is
any object with `matvec(x, *params)` and `vjp(x, lam, *params)` (see
`oracle/operators.py`).  Where the reference leaves a derivative to JAX autodiff
(the cheap wrappers around the custom VJP, `eigh`), the oracle uses the closed form
and the golden vectors in `tests/golden/` (made by running the reference sources)
pin it.

The keyword arguments `symmetric=` / `tridiagonal_cotangent=` of `arnoldi_forward` / `arnoldi_adjoint` are the one
thing here that is NOT the reference: they restate the library's symmetric loops (default off), so that
the shortcut can be checked against the reference restatement and the goldens without a GPU.  Every parity test
of the CUDA path compares with the reference restatement (the defaults).
"""

from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# Arnoldi / Hessenberg  (arnoldi.py)
# ---------------------------------------------------------------------------


def check_reortho_arnoldi(reortho):
    """`arnoldi.py:16-19`: TypeError for anything but "none"/"full"."""
    expected = ["none", "full"]
    if not isinstance(reortho, str) or reortho not in expected:
        raise TypeError(f"Unexpected input for {reortho}: either of {expected} expected.")


def arnoldi_forward(op, krylov_depth, v, *params, reortho_fwd="match", symmetric=False):
    """`arnoldi.py:57-101`.  `reortho_fwd` is what the reference calls `reortho_` inside
    `estimate_backend` (`arnoldi.py:26`): it always equals `reortho_vjp`, whose default
    "match" is != "none", so the second Gram-Schmidt pass runs unless the caller passed
    `reortho_vjp="none"`.

    `symmetric=True` is NOT the reference: it restates the library's `BL_FWD_SYMMETRIC` loop -- the first
    pass takes `h` with columns i-1, i only -- so that the shortcut can be checked against the reference
    restatement (`symmetric=False`) and the goldens on the CPU (tests/test_oracle_golden.py)."""
    v = np.asarray(v)
    n = len(v)
    K = krylov_depth
    if K < 1 or K > n:
        raise ValueError(f"Parameter depth {K} is outside the expected range")
    Q = np.zeros((n, K), dtype=v.dtype)
    H = np.zeros((K, K), dtype=v.dtype)
    length0 = np.sqrt(np.dot(v.conj(), v))
    length = length0
    for i in range(K):
        v = v / length  # arnoldi.py:80
        Q[:, i] = v  # :81
        v = op.matvec(v, *params)  # :84
        h = Q.T.conj() @ v  # :87   (columns > i of Q are zero)
        if symmetric and reortho_fwd != "none":
            h[: max(0, i - 1)] = 0.0  # BL_FWD_SYMMETRIC: q_j^H A q_i = O(eps |A|) for j < i-1
        v = v - Q @ h  # :88
        if reortho_fwd != "none":  # :91
            h2 = Q.T.conj() @ v
            v = v - Q @ h2  # :92  -- h is NOT updated
            if symmetric:  # ... except (library only) for the entries the local first pass skipped: q_j^H v' there
                h[: max(0, i - 1)] = h2[: max(0, i - 1)]  # IS q_j^H A q_i up to rounding, and completes column i of H
        length = np.sqrt(np.dot(v.conj(), v))  # :95
        if i + 1 < K:  # :98  out-of-bounds write at i+1 == K is dropped
            h[i + 1] = length
        H[:, i] = h  # :99
    return Q, H, v, 1.0 / length0


def _lower(m):
    t = np.tril(m)
    return t - 0.5 * np.diag(np.diag(t))


def arnoldi_adjoint(op, params, *, Q, H, r, c, dQ, dH, dr, dc, reortho, symmetric=False, tridiagonal_cotangent=False):
    """`arnoldi.py:104-220`.  Returns `(dv, dparams_tuple)`.

    `symmetric` / `tridiagonal_cotangent` are NOT the reference: they restate the library's `BL_ADJ_SYMMETRIC`
    (`Lambda beta_plus` keeps its super-diagonal term) and `BL_ADJ_TRIDIAG_COTANGENT` (no dQ: `Gamma[idx, j] = 0` for
    `j < idx-2`) loops for CPU checks against the reference restatement."""
    n, K = Q.shape
    dt = Q.dtype
    e1 = np.zeros(K, dtype=dt)
    e1[0] = 1.0
    lower_mask = _lower(np.ones((K, K), dtype=dt))  # :116

    eta = dH[:, K - 1] - Q.T @ dr  # :119
    lam = dr + Q @ eta  # :120
    Lambda = np.zeros_like(Q)
    Gamma = np.zeros((K, K), dtype=dt)
    dp = [np.zeros_like(np.asarray(p)) for p in params]

    Pi_xi = dQ.T + np.outer(eta, r)  # :126
    Pi_gamma = -dc * c * np.outer(e1, e1) + H @ dH.T - dQ.T @ Q  # :127

    P = Q.T.copy()  # :130
    ps = dH.T
    ps_mask = np.tril(np.ones((K, K), dtype=dt), 1)

    beta_minuses = np.concatenate([np.ones(1, dtype=dt), np.diag(H, -1)])  # :136
    alphas = np.diag(H)
    beta_pluses = H - np.diag(np.diag(H)) - np.diag(np.diag(H, -1), -1)  # :138
    symmetric = symmetric and reortho == "full"
    banded = symmetric and tridiagonal_cotangent and not np.any(dQ)
    if symmetric:
        beta_pluses = np.diag(np.diag(H, 1), 1)

    for idx in range(K - 1, -1, -1):  # scan(reverse=True), :162
        p = ps[idx]
        if reortho == "full":  # :201-204
            P = ps_mask[idx][:, None] * P
            p = ps_mask[idx] * p
            lam = lam - P.T @ (P @ lam) + P.T @ p
        vecmat, dp_inc = op.vjp(Q[:, idx], lam, *params)  # :207-208
        dp = [g + h for g, h in zip(dp, dp_inc)]
        Gamma[idx, :] = lower_mask[idx] * (Pi_gamma[idx] - vecmat @ Q)  # :212-213
        if banded:
            Gamma[idx, : max(0, idx - 2)] = 0.0
        Lambda[:, idx] = lam  # :216
        xi = Pi_xi[idx] + (Gamma + Gamma.T)[idx, :] @ Q.T  # :217
        lam = xi - (alphas[idx] * lam - vecmat) - beta_pluses[idx] @ Lambda.T  # :218
        lam = lam / beta_minuses[idx]  # :219
    return lam * c, tuple(dp)  # :166-168


# ---- the same two loops restricted to the columns that are non-zero at each step -------------------------------
# `arnoldi_forward` / `arnoldi_adjoint` above are the literal restatement: every product runs over the full
# (n, K) arrays like the reference's, including the columns that are still zero (forward) or masked (adjoint).
# At the headline size (n = 1M, K = 100, float64: 0.8 GB per basis) that costs minutes and several GB of
# temporaries, so the parity test at that size uses the two functions below: identical arithmetic with the
# identically-zero terms left out, the basis held as K contiguous rows.  tests/test_oracle_golden.py checks them
# against the literal loops.


def arnoldi_forward_active(op, krylov_depth, v, *params, reortho_fwd="match"):
    """`arnoldi.py:57-101` over the active columns only.  Returns `(Qt (K, n), H, r, c)`: `Qt = Q.T`."""
    v = np.asarray(v)
    n, K = len(v), krylov_depth
    if K < 1 or K > n:
        raise ValueError(f"Parameter depth {K} is outside the expected range")
    Qt = np.zeros((K, n), dtype=v.dtype)
    H = np.zeros((K, K), dtype=v.dtype)
    length0 = np.sqrt(np.dot(v, v))
    length = length0
    for i in range(K):
        v = v / length  # :80
        Qt[i] = v  # :81
        v = op.matvec(v, *params)  # :84
        A = Qt[: i + 1]
        h = A @ v  # :87  (the other entries of Q^T v are exact zeros)
        v = v - h @ A  # :88
        if reortho_fwd != "none":  # :91
            v = v - (A @ v) @ A  # :92
        length = np.sqrt(np.dot(v, v))  # :95
        H[: i + 1, i] = h  # :99
        if i + 1 < K:  # :98
            H[i + 1, i] = length
    return Qt, H, v, 1.0 / length0


def arnoldi_adjoint_active(op, params, *, Qt, H, r, c, dQt, dH, dr, dc, reortho):
    """`arnoldi.py:104-220` over the non-zero rows / columns only.  `Qt`, `dQt` are `(K, n)` (`dQt` may be None for
    a zero cotangent, likewise `dr`).  Returns `(dv, dparams_tuple)`."""
    K, n = Qt.shape
    dt = Qt.dtype
    dr_ = np.zeros(n, dtype=dt) if dr is None else dr
    eta = dH[:, K - 1] - (Qt @ dr_ if dr is not None else 0.0)  # :119
    lam = dr_ + eta @ Qt  # :120
    Lt = np.zeros_like(Qt)  # Lambda^T
    Gamma = np.zeros((K, K), dtype=dt)
    dp = [np.zeros_like(np.asarray(p)) for p in params]
    Pi_gamma = -dc * c * np.outer(np.eye(K, 1, dtype=dt), np.eye(K, 1, dtype=dt)) + H @ dH.T  # :127
    if dQt is not None:
        Pi_gamma = Pi_gamma - dQt @ Qt.T
    beta_minuses = np.concatenate([np.ones(1, dtype=dt), np.diag(H, -1)])  # :136
    alphas = np.diag(H)
    for idx in range(K - 1, -1, -1):
        if reortho == "full":  # :201-204, rows <= idx+1 of P = Q^T
            m = min(idx + 2, K)
            P = Qt[:m]
            lam = lam - (P @ lam) @ P + dH[:m, idx] @ P
        vecmat, dp_inc = op.vjp(Qt[idx], lam, *params)  # :207-208
        dp = [g + h for g, h in zip(dp, dp_inc)]
        g_row = Pi_gamma[idx, : idx + 1] - Qt[: idx + 1] @ vecmat  # :212-213 (lower mask: j <= idx, 1/2 on the diagonal)
        g_row[idx] *= 0.5
        Gamma[idx, : idx + 1] = g_row
        Lt[idx] = lam  # :216
        xi = (Gamma + Gamma.T)[idx, :] @ Qt + eta[idx] * r  # :217, Pi_xi[idx] = dQ[:, idx] + eta[idx] r
        if dQt is not None:
            xi = xi + dQt[idx]
        lam = xi - (alphas[idx] * lam - vecmat)  # :218
        if idx + 1 < K:  # beta_plus: row idx of H above the super... (diagonal and sub-diagonal removed)
            lam = lam - H[idx, idx + 1 :] @ Lt[idx + 1 :]
        lam = lam / beta_minuses[idx]  # :219
    return lam * c, tuple(dp)


def tridiag_full_active(op, krylov_depth, v, *params):
    """`lanczos.tridiag(reortho="full")` (`lanczos.py:152-169`) on the active-column loops: returns
    `(((Qt, (alpha, beta)), (r/||r||, ||r||)), pullback)` with `pullback(((dQt | None, (dalpha, dbeta)), (dq_rem | None,
    dnorm)))`."""
    Qt, H, r, c = arnoldi_forward_active(op, krylov_depth, v, *params)
    T = 0.5 * (H + H.T)
    norm = np.linalg.norm(r)
    out = (Qt, (np.diag(T, 0), np.diag(T, 1))), (r / norm, norm)

    def pullback(cot):
        (dQt, (dalpha, dbeta)), (dq_rem, dnorm) = cot
        K = H.shape[0]
        dH = np.diag(np.asarray(dalpha, dtype=H.dtype))
        if K > 1:
            dH = dH + 0.5 * (np.diag(dbeta, 1) + np.diag(dbeta, -1))
        dr = None
        if dq_rem is not None or dnorm:
            dq = np.zeros_like(r) if dq_rem is None else dq_rem
            dr = dq / norm - r * (np.dot(r, dq) / norm**3) + (dnorm or 0.0) * r / norm
        dv, dp = arnoldi_adjoint_active(op, params, Qt=Qt, H=H, r=r, c=c, dQt=dQt, dH=dH, dr=dr, dc=0.0, reortho="full")
        return (dv, *dp)

    return out, pullback


class Hessenberg:
    """`arnoldi.hessenberg(matvec, K, reortho=, custom_vjp=, reortho_vjp=)` (`arnoldi.py:7-54`).

    `__call__(v, *params)` -> `(Q, H, r, c)`; `vjp(v, *params)` -> `(outputs, pullback)` with
    `pullback((dQ, dH, dr, dc))` -> `(dv, *dparams)` as `estimate_bwd` (`arnoldi.py:33-49`).
    """

    def __init__(self, op, krylov_depth, *, reortho, reortho_vjp="match"):
        check_reortho_arnoldi(reortho)
        self.op, self.K = op, krylov_depth
        self.reortho, self.reortho_vjp = reortho, reortho_vjp

    def __call__(self, v, *params):
        return arnoldi_forward(self.op, self.K, v, *params, reortho_fwd=self.reortho_vjp)

    def vjp(self, v, *params):
        Q, H, r, c = self(v, *params)

        def pullback(cot):
            dQ, dH, dr, dc = cot
            dv, dp = arnoldi_adjoint(
                self.op, params, Q=Q, H=H, r=r, c=c, dQ=dQ, dH=dH, dr=dr, dc=dc,
                reortho=self.reortho,
            )  # fmt: skip
            return (dv, *dp)

        return (Q, H, r, c), pullback


# ---------------------------------------------------------------------------
# Lanczos tridiagonalisation  (lanczos.py)
# ---------------------------------------------------------------------------


def tridiag_from_hessenberg(Q, H, r):
    """`lanczos.py:159-167`: symmetrise, take the diagonals, normalise the remainder."""
    T = 0.5 * (H + H.T)
    norm = np.linalg.norm(r)
    return (Q.T, (np.diag(T, 0), np.diag(T, 1))), (r / norm, norm)


def tridiag_from_hessenberg_pullback(Q, H, r, cot):
    """Cotangent map of `tridiag_from_hessenberg` (JAX autodiff does this in the
    reference).  `cot = ((dQt, (dalpha, dbeta)), (dq_rem, dnorm))` -> `(dQ, dH, dr, dc)`."""
    (dQt, (dalpha, dbeta)), (dq_rem, dnorm) = cot
    K = H.shape[0]
    dH = np.diag(np.asarray(dalpha, dtype=H.dtype))
    if K > 1:
        dH = dH + 0.5 * (np.diag(dbeta, 1) + np.diag(dbeta, -1))
    norm = np.linalg.norm(r)
    # two uses of norm(r): r/norm and norm itself
    dr = dq_rem / norm - r * (np.dot(r, dq_rem) / norm**3) + dnorm * r / norm
    return np.asarray(dQt).T, dH, dr, np.zeros((), dtype=H.dtype)


class TridiagFull:
    """`lanczos.tridiag(matvec, K, reortho="full")` -> `_tridiag_reortho_full`
    (`lanczos.py:152-169`): Arnoldi with `reortho="full"` and symmetrisation."""

    def __init__(self, op, krylov_depth):
        self.alg = Hessenberg(op, krylov_depth, reortho="full")

    def __call__(self, v, *params):
        Q, H, r, _c = self.alg(v, *params)
        return tridiag_from_hessenberg(Q, H, r)

    def vjp(self, v, *params):
        (Q, H, r, _c), pull = self.alg.vjp(v, *params)

        def pullback(cot):
            return pull(tridiag_from_hessenberg_pullback(Q, H, r, cot))

        return tridiag_from_hessenberg(Q, H, r), pullback


def lanczos3_forward(op, krylov_depth, vec, *params):
    """Three-term recurrence without re-orthogonalisation, `lanczos.py:215-285`.
    Returns `(decomposition, remainder, 1/||vec||)`."""
    K = krylov_depth
    vec = np.asarray(vec)
    n = len(vec)
    xs = np.zeros((K + 1, n), dtype=vec.dtype)
    a = np.zeros(K, dtype=vec.dtype)
    b = np.zeros(K, dtype=vec.dtype)
    x_prev = None
    x = vec / np.linalg.norm(vec)  # :222
    xs[0] = x
    b_prev = 0.0
    for i in range(K):
        ax = op.matvec(x, *params)
        a[i] = x @ ax  # :256 / :280
        res = ax - a[i] * x
        if i > 0:
            res = res - b_prev * x_prev  # :282
        b[i] = np.linalg.norm(res)  # :283
        x_prev, x, b_prev = x, res / b[i], b[i]
        xs[i + 1] = x
    return (xs[:-1], (a, b[:-1])), (xs[-1], b[-1]), 1.0 / np.linalg.norm(vec)


def lanczos3_adjoint(op, params, *, initvec_norm, alphas, betas, xs, dalphas, dbetas, dxs):
    """`lanczos.py:288-335`.  `xs`, `dxs` are `(K+1, n)`; `betas`, `dbetas` length `K`.
    Returns `(grad_initvec, grad_param)`; the reference supports exactly one parameter
    (`lanczos.py:329`)."""
    K = len(alphas)
    xi = -dxs[K]  # :303
    lam_plus = np.zeros_like(xi)
    grad = None
    for k in range(K - 1, -1, -1):
        x, xplus = xs[k], xs[k + 1]  # "xs": (xs[1:], xs[:-1]), :298
        xi = xi / betas[k]  # :322
        mu = dbetas[k] - lam_plus @ x + xplus @ xi  # :323
        nu = dalphas[k] + x @ xi  # :324
        lam = -xi + mu * xplus + nu * x  # :325
        # matvec applied to lambda, cotangent x  (:328-329)
        _, (inc,) = op.vjp(lam, x, *params)
        a_lam = op.matvec(lam, *params)
        grad = inc if grad is None else grad + inc
        xi = -dxs[k] - a_lam + alphas[k] * lam + betas[k] * lam_plus - betas[k] * nu * xplus  # :332
        lam_plus = lam
    # the carry unpacked as `lambda_1` in the reference is the final xi
    grad_initvec = ((xi @ xs[0]) * xs[0] - xi) / initvec_norm  # :311
    return grad_initvec, grad


class TridiagNone:
    """`lanczos.tridiag(matvec, K, reortho="none")` -> `_tridiag_reortho_none`
    (`lanczos.py:172-212`)."""

    def __init__(self, op, krylov_depth):
        self.op, self.K = op, krylov_depth

    def __call__(self, v, *params):
        dec, rem, _ = lanczos3_forward(self.op, self.K, v, *params)
        return dec, rem

    def vjp(self, v, *params):
        (xs, (a, b)), (x_last, b_last) = out = self(v, *params)
        vnorm = np.linalg.norm(v)

        def pullback(cot):
            (dxs, (da, db)), (dx_last, db_last) = cot
            g_v, g_p = lanczos3_adjoint(
                self.op, params, initvec_norm=vnorm, alphas=a,
                betas=np.concatenate([b, [b_last]]),
                xs=np.concatenate([xs, x_last[None]]),
                dalphas=da, dbetas=np.concatenate([db, [db_last]]),
                dxs=np.concatenate([dxs, dx_last[None]]),
            )  # fmt: skip
            return g_v, g_p

        return out, pullback


def tridiag(op, krylov_depth, *, reortho):
    """`lanczos.tridiag` dispatch and error type (`lanczos.py:142-149`)."""
    if reortho == "full":
        return TridiagFull(op, krylov_depth)
    if reortho == "none":
        return TridiagNone(op, krylov_depth)
    raise ValueError(f"reortho={reortho} unsupported. Choose eiter {'full', 'none'}.")


# ---------------------------------------------------------------------------
# SLQ integrand  (lanczos.py:14-139)
# ---------------------------------------------------------------------------


def dense_tridiag(alpha, beta):
    return np.diag(alpha) + np.diag(beta, 1) + np.diag(beta, -1)


def quadform_of_tridiag(matfun, alpha, beta):
    """`e1^T f(T) e1` via `eigh` (`lanczos.py:48-59`)."""
    w, U = np.linalg.eigh(dense_tridiag(alpha, beta))
    return np.dot(U[0], matfun(w) * U[0]), (w, U)


def quadform_of_tridiag_grad(matfun, matfun_grad, w, U):
    """Closed-form cotangents `(dalpha, dbeta)` of `e1^T f(T) e1` (Daleckii-Krein);
    the reference differentiates through `eigh` with JAX (`lanczos.py:53-59`)."""
    fw, dfw = matfun(w), matfun_grad(w)
    dw = w[:, None] - w[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        F = (fw[:, None] - fw[None, :]) / dw
    same = np.abs(dw) <= 1e-14 * np.maximum(1.0, np.abs(w).max())
    F[same] = (0.5 * (dfw[:, None] + dfw[None, :]))[same]
    G = U @ (np.outer(U[0], U[0]) * F) @ U.T
    return np.diag(G), np.diag(G, 1) + np.diag(G, -1)


class IntegrandSPD:
    """`lanczos.integrand_spd(matfun, K, matvec, reortho=)` (`lanczos.py:14-61`).

    `__call__(v0, *params)` -> scalar; `value_and_grad(v0, *params)` ->
    `(value, (dv0, *dparams))`.  `matfun_grad` is `f'` (JAX derives it in the reference)."""

    def __init__(self, matfun, matfun_grad, krylov_depth, op, *, reortho="full"):
        self.f, self.df = matfun, matfun_grad
        self.alg = tridiag(op, krylov_depth, reortho=reortho)

    def __call__(self, v0, *params):
        v0 = np.asarray(v0).ravel()
        scale = np.linalg.norm(v0)  # :25
        (_, (alpha, beta)), _ = self.alg(v0 / scale, *params)
        val, _ = quadform_of_tridiag(self.f, alpha, beta)
        return scale**2 * val  # :59

    def value_and_grad(self, v0, *params):
        v0 = np.asarray(v0).ravel()
        scale = np.linalg.norm(v0)
        u = v0 / scale
        ((Qt, (alpha, beta)), (q_rem, _)), pull = self.alg.vjp(u, *params)
        g, (w, U) = quadform_of_tridiag(self.f, alpha, beta)
        dalpha, dbeta = quadform_of_tridiag_grad(self.f, self.df, w, U)
        cot = (
            (np.zeros_like(Qt), (scale**2 * dalpha, scale**2 * dbeta)),
            (np.zeros_like(q_rem), np.zeros((), dtype=Qt.dtype)),
        )
        du, *dparams = pull(cot)
        # chain rule through u = v0/||v0|| and the scale**2 factor
        dv0 = 2.0 * g * v0 + (du - u * np.dot(u, du)) / scale
        return scale**2 * g, (dv0, *dparams)


class IntegrandSPDReuse:
    """`lanczos.integrand_spd_custom_vjp_reuse` (`lanczos.py:64-139`): same value, cheap
    inexact parameter gradient `d/dtheta <w1, A(w2; theta)>` and `dv0 := 0` (`:130-134`)."""

    def __init__(self, matfun, matfun_grad, order, op, *, reortho="full"):
        self.f, self.df, self.op = matfun, matfun_grad, op
        self.alg = tridiag(op, order, reortho=reortho)

    def value_and_grad(self, v0, *params):
        v0 = np.asarray(v0).ravel()
        scale = np.linalg.norm(v0)
        u = v0 / scale
        (Qt, (alpha, beta)), _ = self.alg(u, *params)
        val, (w, U) = quadform_of_tridiag(self.f, alpha, beta)
        sol = U @ (self.df(w) * U[0])  # :112-113
        w1, w2 = scale**2 * (Qt.T @ sol), u  # :114
        _, dparams = self.op.vjp(w2, w1, *params)  # :121
        return scale**2 * val, (np.zeros_like(v0), *dparams)

    def __call__(self, v0, *params):
        return self.value_and_grad(v0, *params)[0]


# ---------------------------------------------------------------------------
# Hutchinson  (hutchinson.py:51-54 and matfree's hutchinson.hutchinson)
# ---------------------------------------------------------------------------


def hutchinson_mean(integrand, samples, *params):
    """Mean over explicitly supplied probes `samples (num, n)` (`hutchinson.py:51-54`)."""
    vals = [integrand(s, *params) for s in samples]
    return np.mean(vals, axis=0)


def hutchinson_value_and_grad(integrand, samples, *params):
    """Value and parameter gradient of `hutchinson_mean`; probes are constants
    (`stop_gradient`, `hutchinson.py:12`)."""
    val, grads = 0.0, None
    for s in samples:
        v, (_dv0, *dp) = integrand.value_and_grad(s, *params)
        val = val + v
        grads = dp if grads is None else [g + h for g, h in zip(grads, dp)]
    num = len(samples)
    return val / num, tuple(g / num for g in grads)


# ---------------------------------------------------------------------------
# Matrix-exponential action (util/pde_util.py:257-268)
# ---------------------------------------------------------------------------


def expm_action(op, krylov_depth, dt, y0, *params, reortho="full"):
    """`expm_arnoldi.expm` (`/root/reference/src/matfree_extensions/util/pde_util.py:260-266`):
    `(1/c) Q expm(dt H) e1`."""
    import scipy.linalg

    Q, H, _r, c = Hessenberg(op, krylov_depth, reortho=reortho)(y0, *params)
    E = scipy.linalg.expm(dt * H)
    return (1.0 / c) * (Q @ E[:, 0])


def expm_action_vjp(op, krylov_depth, dt, y0, params, cot, reortho="full"):
    """VJP of `expm_action` for an output cotangent `cot (n,)`: the `expm` Fréchet
    adjoint (JAX autodiff in the reference) feeding the Arnoldi adjoint."""
    import scipy.linalg

    alg = Hessenberg(op, krylov_depth, reortho=reortho)
    (Q, H, r, c), pull = alg.vjp(y0, *params)
    K = H.shape[0]
    E = scipy.linalg.expm(dt * H)
    y = E[:, 0]
    # out = (1/c) Q y ;  y = expm(dt H) e1
    dQ = np.outer(cot, y) / c
    dy = (Q.T @ cot) / c
    dc = -np.dot(cot, Q @ y) / c**2
    # d<dy, expm(dt H) e1>/dH = dt * L_expm(dt H^T)[dy e1^T]
    e1 = np.zeros(K)
    e1[0] = 1.0
    dH = dt * scipy.linalg.expm_frechet(dt * H.T, np.outer(dy, e1), compute_expm=False)
    return pull((dQ, dH, np.zeros_like(r), dc))
