"""Synthetic operands of the shapes BASELINE.json names (no datasets, no network)."""

from __future__ import annotations

import numpy as np


def banded_spd_coo(n: int, bands: int = 5, seed: int = 0, max_offset: int = 2000, long_range: int = 1):
    """Symmetric, strictly diagonally dominant (=> SPD) sparse operand in the COO layout
    `suite_sparse_load` produces for a symmetric MatrixMarket file: diagonal + strict lower
    triangle in "file order", then the mirrored strict upper triangle
    (`/root/reference/src/matfree_extensions/util/exp_util.py:35-42`).

    `bands` sub-diagonals at random offsets in [1, max_offset] (FEM-like bandwidth), of which
    `long_range` sit at offsets ~ n/3 (SuiteSparse-like long-range coupling); off-diagonal values
    U(-1, 0), diagonal 2*bands + 2: spectrum within about (1, 4*bands + 3), so Lanczos does not
    break down through depth 100.  Rows hold 2*bands + 1 entries (11 for the default).
    Returns `(row int32, col int32, data float64)`.
    """
    rng = np.random.default_rng(seed)
    hi = max(1, min(max_offset, n - 1))
    near = rng.choice(np.arange(1, hi + 1), size=min(bands - long_range, hi), replace=False)
    far = (n // 3 + rng.integers(0, max(1, n // 7), size=long_range)) if n > 8 * max_offset else np.array([], int)
    offs = np.unique(np.concatenate([near, far]).astype(np.int64))
    offs = offs[(offs >= 1) & (offs < n)]
    lo_r = np.concatenate([np.arange(o, n, dtype=np.int32) for o in offs]) if len(offs) else np.zeros(0, np.int32)
    lo_c = np.concatenate([np.arange(0, n - o, dtype=np.int32) for o in offs]) if len(offs) else np.zeros(0, np.int32)
    vals = -rng.uniform(0.0, 1.0, lo_r.size)
    diag = np.arange(n, dtype=np.int32)
    row = np.concatenate([diag, lo_r, lo_c])
    col = np.concatenate([diag, lo_c, lo_r])
    data = np.concatenate([np.full(n, 2.0 * len(offs) + 2.0), vals, vals])
    return row, col, data


def slq_cotangent_dH(dalpha, dbeta, dtype):
    """Cotangent of `H` for cotangents `(dalpha, dbeta)` on the tridiagonal coefficients
    (`T = (H + H^T)/2`, `lanczos.py:162-164`): `diag(dalpha) + (superdiag + subdiag)(dbeta)/2`."""
    dalpha = np.asarray(dalpha, dtype=dtype)
    dH = np.diag(dalpha)
    if len(dalpha) > 1:
        dbeta = np.asarray(dbeta, dtype=dtype)
        dH = dH + 0.5 * (np.diag(dbeta, 1) + np.diag(dbeta, -1))
    return np.ascontiguousarray(dH, dtype=dtype)
