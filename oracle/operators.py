"""Oracle matvec backends (NumPy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Every operator is a pair

    matvec(x, *params)      -> A(x; params)
    vjp(x, lam, *params)    -> (A^T lam, (d<lam, A(x;p)>/dp for p in params))

which is what `jax.vjp(lambda u, p: matvec(u, *p), q, params)` hands the
reference's adjoint sweep (`/root/reference/src/matfree_extensions/arnoldi.py:207-209`).
"""

from __future__ import annotations

import numpy as np


class DenseOperator:
    """`matvec = lambda s, p: p @ s` — the operator of the reference's Arnoldi tests
    (`/root/reference/tests/test_arnoldi/test_hessenberg_forward.py:20`,
    `test_hessenberg_adjoint.py:22-26`)."""

    num_params = 1

    def matvec(self, x, p):
        return p @ x

    def vjp(self, x, lam, p):
        return p.T @ lam, (np.outer(lam, x),)


class SymDenseOperator:
    """`matvec = lambda s, p: (p + p.T) @ s` — the operator of
    `/root/reference/tests/test_lanczos/test_tridiag_adjoint.py:20-21` and
    `test_hessenberg_adjoint.py:65-66` (BASELINE config 1)."""

    num_params = 1

    def matvec(self, x, p):
        return (p + p.T) @ x

    def vjp(self, x, lam, p):
        return (p + p.T) @ lam, (np.outer(lam, x) + np.outer(x, lam),)


class CooOperator:
    """Sparse operator in the reference's BCOO convention
    (`/root/reference/experiments/benchmarks/wall_times_vjp_through_lanczos_arnoldi/suite_sparse/benchmark.py:61-68`,
    `/root/reference/src/matfree_extensions/util/exp_util.py:35-42`):
    the parameter vector is the COO `data` array, in COO order; indices are a
    fixed `(nnz, 2)` int32 array; duplicates are summed by the matvec and every
    stored entry is an independent parameter.
    """

    num_params = 1

    def __init__(self, row, col, shape):
        self.row = np.asarray(row, dtype=np.int32)
        self.col = np.asarray(col, dtype=np.int32)
        self.shape = tuple(shape)

    def matvec(self, x, p):
        y = np.zeros(self.shape[0], dtype=np.result_type(x, p))
        np.add.at(y, self.row, p * x[self.col])
        return y

    def vjp(self, x, lam, p):
        xbar = np.zeros(self.shape[1], dtype=np.result_type(lam, p))
        np.add.at(xbar, self.col, p * lam[self.row])
        return xbar, (lam[self.row] * x[self.col],)


class CsrFastOperator:
    """Same arithmetic as `CooOperator`, evaluated with SciPy CSR products so that
    the CPU baseline of `bench.py` runs at a realistic speed.  `perm` maps CSR
    slots back to COO parameter positions; duplicates stay separate slots.
    Integer work follows `coo_to_csr` below (bit-exact contract)."""

    num_params = 1

    def __init__(self, row, col, shape):
        import scipy.sparse as sp

        self.shape = tuple(shape)
        self.row = np.asarray(row, dtype=np.int32)
        self.col = np.asarray(col, dtype=np.int32)
        self.row_ptr, self.col_idx, self.perm = coo_to_csr(self.row, self.col, shape[0])
        self._sp = sp
        self._rows_csr = self.row[self.perm]

    def _matrix(self, p):
        return self._sp.csr_matrix(
            (p[self.perm], self.col_idx, self.row_ptr), shape=self.shape
        )

    def matvec(self, x, p):
        return self._matrix(p) @ x

    def vjp(self, x, lam, p):
        xbar = self._matrix(p).T @ lam
        return xbar, (lam[self.row] * x[self.col],)


def coo_to_csr(row, col, nrows):
    """COO -> CSR index work (bit-exact contract, SURVEY §8a `suite_sparse_load` row).

    Stable sort of the COO entries by (row, col); entries with equal (row, col)
    keep their COO order and stay separate slots.  Returns
    `(row_ptr int32[nrows+1], col_idx int32[nnz], perm int32[nnz])` with
    `perm[k]` = COO position of CSR slot `k`.
    """
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    perm = np.lexsort((col, row)).astype(np.int32)  # lexsort is stable
    counts = np.bincount(row, minlength=nrows)
    row_ptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    return row_ptr.astype(np.int32), col[perm].astype(np.int32), perm


def csr_to_sell(row_ptr, nrows, chunk=32):
    """CSR -> SELL-C (sliced ELLPACK, slice height `chunk`, no row sorting) index work.

    Slice `s` covers rows `[s*chunk, (s+1)*chunk)`, has width
    `w_s = max row length in the slice`, and occupies `w_s*chunk` slots starting at
    `slice_ptr[s]`; slot of (row r, k-th entry) = `slice_ptr[s] + k*chunk + r%chunk`.
    Returns `(slice_ptr int64[nslices+1], slot_of_csr int64[nnz])`; padded slots
    are the ones no CSR entry maps to.
    """
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    lens = np.diff(row_ptr)
    nslices = (nrows + chunk - 1) // chunk
    padded = np.zeros(nslices * chunk, dtype=np.int64)
    padded[:nrows] = lens
    widths = padded.reshape(nslices, chunk).max(axis=1)
    slice_ptr = np.zeros(nslices + 1, dtype=np.int64)
    np.cumsum(widths * chunk, out=slice_ptr[1:])
    nnz = int(row_ptr[-1])
    rows_of = np.repeat(np.arange(nrows, dtype=np.int64), lens)
    k_of = np.arange(nnz, dtype=np.int64) - row_ptr[rows_of]
    slot = slice_ptr[rows_of // chunk] + k_of * chunk + rows_of % chunk
    return slice_ptr, slot


def mm_expand_symmetric(row, col, data):
    """What `scipy.io.mmread` does for a `symmetric` MatrixMarket file and hence the
    parameter order `suite_sparse_load` produces
    (`/root/reference/src/matfree_extensions/util/exp_util.py:35-42`): stored
    lower-triangular entries in file order, followed by the mirrored strictly
    off-diagonal entries in the same order."""
    row = np.asarray(row)
    col = np.asarray(col)
    data = np.asarray(data)
    off = row != col
    return (
        np.concatenate([row, col[off]]),
        np.concatenate([col, row[off]]),
        np.concatenate([data, data[off]]),
    )


# ---------------------------------------------------------------------------
# Gaussian-process kernels (reference: util/gp_util.py)
# ---------------------------------------------------------------------------


def softplus(x, beta=1.0, threshold=20.0):
    """`/root/reference/src/matfree_extensions/util/gp_util.py:188-199`: identity above
    the threshold, `log(1+exp(beta x))/beta` below."""
    x = np.asarray(x)
    safe = np.where(x * beta < threshold, x, np.ones_like(x))
    return np.where(x * beta < threshold, np.log1p(np.exp(beta * safe)) / beta, x)


def softplus_grad(x, beta=1.0, threshold=20.0):
    x = np.asarray(x)
    safe = np.where(x * beta < threshold, x, np.ones_like(x))
    return np.where(x * beta < threshold, 1.0 / (1.0 + np.exp(-beta * safe)), 1.0)


class GramOperator:
    """`(K(X,X) + noise I) v` with a scaled Matérn-3/2, Matérn-1/2 or RBF kernel.

    Kernels: `/root/reference/src/matfree_extensions/util/gp_util.py:69-107` (Matérn-3/2),
    `:110-148` (Matérn-1/2), `:151-184` (RBF); Gram matvec `:525-543`; noise term as in
    `likelihood_pdf_p.cov_matvec` (`:252-268`).  Parameters, in order:
    `raw_lengthscale (d,)`, `raw_outputscale ()`, `noise ()` — `noise` is the already
    constrained value added to the diagonal (the caller applies its own constraint,
    `/root/reference/experiments/applications/gaussian_process/train/optim_logml_adjoints_adaptive.py:66,124`).
    The Gram matrix is evaluated in row blocks so the oracle runs at n ~ 10^4.
    """

    num_params = 3

    def __init__(self, X, kind="matern32", block=2048):
        self.X = np.asarray(X)
        self.kind = kind
        self.block = block
        if kind not in ("matern32", "matern12", "rbf"):
            raise ValueError(kind)

    def _scaled(self, raw_ls):
        ls = softplus(raw_ls)
        fac = np.sqrt(np.asarray(3.0, dtype=self.X.dtype)) if self.kind == "matern32" else 1.0
        return (fac * self.X / ls).astype(self.X.dtype), ls, fac

    def _block(self, Xi, Xs, sigma):
        dt = Xs.dtype
        xx = np.einsum("id,id->i", Xi, Xi)
        yy = np.einsum("jd,jd->j", Xs, Xs)
        s2 = np.maximum(0.0, xx[:, None] + yy[None, :] - 2.0 * (Xi @ Xs.T))
        if self.kind == "rbf":
            k = sigma * np.exp(-s2 / 2)
            return k, s2, None
        s = np.sqrt(s2 + np.finfo(dt).eps)
        if self.kind == "matern32":
            return sigma * (1 + s) * np.exp(-s), s2, s
        return sigma * np.exp(-s), s2, s

    def matvec(self, v, raw_ls, raw_os, noise):
        Xs, _, _ = self._scaled(raw_ls)
        sigma = softplus(raw_os)
        out = np.empty_like(v)
        for i0 in range(0, len(v), self.block):
            k, _, _ = self._block(Xs[i0 : i0 + self.block], Xs, sigma)
            out[i0 : i0 + self.block] = k @ v
        return out + noise * v

    def vjp(self, q, lam, raw_ls, raw_os, noise):
        """Returns `(K^T lam + noise lam, (d raw_ls, d raw_os, d noise))` for the scalar
        `<lam, (K + noise I) q>`; hand-derived chain rule through the kernel, the clamp
        (`max(0, .)` has zero derivative where it clamps), and the soft-plus."""
        Xs, ls, fac = self._scaled(raw_ls)
        sigma = softplus(raw_os)
        n, d = Xs.shape
        xbar = np.zeros_like(q)
        d_sigma = 0.0
        d_Xs = np.zeros_like(Xs)  # gradient w.r.t. the scaled inputs
        for i0 in range(0, n, self.block):
            Xi = Xs[i0 : i0 + self.block]
            li = lam[i0 : i0 + self.block]
            k, s2, s = self._block(Xi, Xs, sigma)
            xbar += k.T @ li
            w = li[:, None] * q[None, :]  # weight of k_ij in the scalar
            d_sigma += np.sum(w * k) / sigma
            if self.kind == "rbf":
                dk_ds2 = -0.5 * k
            elif self.kind == "matern32":
                # k = sigma (1+s) e^{-s}; dk/ds = -sigma s e^{-s}; ds/ds2 = 1/(2 s)
                dk_ds2 = -0.5 * sigma * np.exp(-s)
            else:
                dk_ds2 = -k / (2 * s)
            g = w * dk_ds2 * (s2 > 0)  # d scalar / d s2_ij
            # s2_ij = |xi|^2 + |xj|^2 - 2 xi.xj  (as expanded in the reference)
            d_Xs[i0 : i0 + self.block] += 2 * (g.sum(1)[:, None] * Xi - g @ Xs)
            d_Xs += 2 * (g.sum(0)[:, None] * Xs - g.T @ Xi)
        # Xs = fac * X / ls  ->  d/d ls_k = -sum_i dXs_ik * Xs_ik / ls_k
        d_ls = -(d_Xs * Xs).sum(0) / ls
        d_raw_ls = d_ls * softplus_grad(raw_ls)
        d_raw_os = d_sigma * softplus_grad(raw_os)
        return xbar + noise * lam, (d_raw_ls, d_raw_os, np.dot(lam, q))


# ---------------------------------------------------------------------------
# Wave-equation stencil operator (reference: util/pde_util.py)
# ---------------------------------------------------------------------------


class WaveStencilOperator:
    """Linear right-hand side of the anisotropic wave equation on a `g x g` grid:
    state `(u, du)` raveled to `2 g^2`, `A(u,du) = (du, scale^2 * conv(stencil, pad(u)))`.

    Follows `pde_wave_anisotropic.rhs` (`/root/reference/src/matfree_extensions/util/pde_util.py:126-143`)
    with `boundary_neumann` edge-replicate padding (`:153-157`), a 3x3 stencil applied
    as a true convolution (`jax.scipy.signal.convolve2d(stencil, padded, "valid")`, `:137`)
    and `constrain=jnp.square`
    (`/root/reference/experiments/applications/partial_differential_equation/train.py:57-59`).
    `stencil_laplacian` reproduces `pde_util.py:18-20` literally: the centre weight is
    -2 (not -4) — a quirk of the reference that parity keeps.
    Parameter: `scale (g, g)` (raveled `(g*g,)` accepted).
    """

    num_params = 1

    @staticmethod
    def stencil_laplacian(dx):
        return np.asarray([[0.0, 1.0, 0.0], [1.0, -2.0, 1.0], [0.0, 1.0, 0.0]]) / dx**2

    def __init__(self, grid, stencil):
        self.g = int(grid)
        self.stencil = np.asarray(stencil, dtype=np.float64)
        assert self.stencil.shape == (3, 3)

    def _conv(self, u):
        g = self.g
        up = np.pad(u, 1, mode="edge")
        out = np.zeros_like(u)
        for a in range(3):
            for b in range(3):
                wgt = self.stencil[a, b]
                if wgt != 0.0:
                    out += wgt.astype(u.dtype) * up[2 - a : 2 - a + g, 2 - b : 2 - b + g]
        return out

    def _conv_T(self, w):
        # transpose of `_conv` (scatter into the padded grid, then fold the halo back
        # onto the edge cells = transpose of np.pad(mode="edge"))
        g = self.g
        pad = np.zeros((g + 2, g + 2), dtype=w.dtype)
        for a in range(3):
            for b in range(3):
                wgt = self.stencil[a, b]
                if wgt != 0.0:
                    pad[2 - a : 2 - a + g, 2 - b : 2 - b + g] += wgt.astype(w.dtype) * w
        pad[1, :] += pad[0, :]
        pad[-2, :] += pad[-1, :]
        pad[:, 1] += pad[:, 0]
        pad[:, -2] += pad[:, -1]
        return pad[1:-1, 1:-1].copy()

    def matvec(self, x, scale):
        g = self.g
        u, du = x[: g * g].reshape(g, g), x[g * g :].reshape(g, g)
        s2 = np.square(np.asarray(scale).reshape(g, g))
        return np.concatenate([du.ravel(), (s2 * self._conv(u)).ravel()])

    def vjp(self, x, lam, scale):
        g = self.g
        u = x[: g * g].reshape(g, g)
        lu, ldu = lam[: g * g].reshape(g, g), lam[g * g :].reshape(g, g)
        sc = np.asarray(scale).reshape(g, g)
        # y_u = du ; y_du = sc^2 * conv(u)
        xbar_u = self._conv_T(sc * sc * ldu)
        xbar_du = lu
        dscale = 2.0 * sc * ldu * self._conv(u)
        return (
            np.concatenate([xbar_u.ravel(), xbar_du.ravel()]),
            (dscale.reshape(np.shape(scale)),),
        )
