"""Low-rank approximations: drop-ins for `/root/reference/src/matfree_extensions/low_rank.py`.

    cholesky = low_rank.cholesky_partial_pivot(rank=100)
    precondition = low_rank.preconditioner(cholesky)
    pre, info = precondition(lazy_kernel, n)      # lazy_kernel = operators.bound(gram_op, raw_ls, raw_os, noise)
    z = pre(v, noise)                             # (noise I + L L^T)^{-1} v
    x, _ = cg.pcg_fixed_step(50)(A, b, pre.bind(noise))

The lazily evaluated matrix is an operator object that exposes its elements (dense operand: its matrix;
Gram operand: the kernel matrix WITHOUT the noise term, as `likelihood_pdf_p.lazy_kernel`,
`gp_util.py:257-258`).  The factorisation (`bl_cholesky_partial`) keeps pivots, permutation and the factor
on the device.  As in the reference nothing here is differentiable.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from experiments_lanczos_adjoints_b200 import _lib
from experiments_lanczos_adjoints_b200 import device as dev
from experiments_lanczos_adjoints_b200.operators import BoundOperator


def _factorise(lazy_kernel, n, rank, pivot, dtype, stream):
    if not isinstance(lazy_kernel, BoundOperator):
        raise TypeError("lazy_kernel must be operators.bound(op, *params) with an element-providing operator")
    if rank > n:  # low_rank.py:67-69
        raise ValueError(f"Rank exceeds n: {rank} >= {n}.")
    if rank < 1:  # low_rank.py:70-72
        raise ValueError(f"Rank must be positive, but {rank} < {1}.")
    stream = stream or dev.default_stream()
    bound = lazy_kernel.bind(dtype, stream)
    dtype = np.dtype(dtype)
    L = dev.DeviceArray((rank, n), dtype, ld=dev.basis_ld(n, dtype))
    nbytes = _lib.load().bl_cholesky_workspace_bytes(n, rank, dev.dtype_code(dtype))
    ws = dev.DeviceArray(((nbytes + 7) // 8,), np.float64)
    ok = C.c_int(1)
    pivots = (C.c_int64 * rank)()
    _lib.call("bl_cholesky_partial", lazy_kernel.op._handle, dev.dtype_code(dtype), n, rank, int(pivot), L.ptr, L.ld,
              C.byref(ok), pivots, ws.ptr, nbytes, stream.ptr)  # fmt: skip
    del bound
    return L, bool(ok.value), np.asarray(list(pivots), dtype=np.int64)


def _dtype_of(lazy_kernel, dtype):
    if not isinstance(lazy_kernel, BoundOperator):
        raise TypeError("lazy_kernel must be operators.bound(op, *params) with an element-providing operator")
    if dtype is not None:
        return np.dtype(dtype)
    for p in lazy_kernel.params:
        if hasattr(p, "dtype") and np.dtype(p.dtype).kind == "f":
            return np.dtype(p.dtype) if np.dtype(p.dtype).itemsize >= 4 else np.dtype(np.float32)
    return np.dtype(np.float64)


def cholesky_partial(*, rank: int, dtype=None):
    """`low_rank.cholesky_partial` (`low_rank.py:63-118`): `cholesky(lazy_kernel, n) -> (L (n, rank), {})`."""

    def cholesky(lazy_kernel, n: int, /, *, stream=None):
        L, _, _ = _factorise(lazy_kernel, n, rank, False, _dtype_of(lazy_kernel, dtype), stream)
        return L.T, {}

    return cholesky


def cholesky_partial_pivot(*, rank: int, dtype=None):
    """`low_rank.cholesky_partial_pivot` (`low_rank.py:120-225`): `-> (L (n, rank), {"success": bool})`;
    the info also carries the chosen `pivots` (original indices)."""

    def cholesky(matrix_element, n: int, /, *, stream=None):
        L, ok, pivots = _factorise(matrix_element, n, rank, True, _dtype_of(matrix_element, dtype), stream)
        return L.T, {"success": ok, "pivots": pivots}

    return cholesky


class _Preconditioner:
    """`solve(v, s)` of `low_rank.py:30-47`; `bind(s)` fixes the shift for use inside PCG."""

    def __init__(self, chol_t: dev.DeviceArray, stream):
        self._rows = chol_t.T  # storage (rank, n)
        rank, n = self._rows._shape
        if rank > n:  # low_rank.py:27-28: tall, not wide
            raise AssertionError((n, rank))
        self.n, self.rank, self.dtype = n, rank, self._rows.dtype
        h = C.c_void_p()
        _lib.call("bl_precond_create", dev.dtype_code(self.dtype), n, rank, self._rows.ptr, self._rows.ld, stream.ptr,
                  C.byref(h))  # fmt: skip
        self._handle, self._shift = h.value, None

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _lib.load().bl_precond_destroy(h)
            except Exception:
                pass

    def _precond_handle(self):
        return self._handle

    def bind(self, s, stream=None):
        s = float(np.asarray(s if not isinstance(s, dev.DeviceArray) else s.numpy()).reshape(-1)[0])
        if s != self._shift:
            _lib.call("bl_precond_set_shift", self._handle, s, (stream or dev.default_stream()).ptr)
            self._shift = s
        return self

    def __call__(self, v, s, *, stream=None):
        stream = stream or dev.default_stream()
        self.bind(s, stream)
        v = dev.asarray(v, dtype=self.dtype)
        out = dev.DeviceArray((self.n,), self.dtype)
        _lib.call("bl_precond_apply", self._handle, dev.dtype_code(self.dtype), v.ptr, out.ptr, stream.ptr)
        return out


def preconditioner(cholesky, /):
    """`low_rank.preconditioner` (`low_rank.py:10-60`): turn a low-rank factorisation into
    `v, s -> (s I + L L^T)^{-1} v`."""

    def solve_with_preconditioner(lazy_kernel, /, nrows: int, *, stream=None):
        stream = stream or dev.default_stream()
        chol, info = cholesky(lazy_kernel, nrows, stream=stream)
        return _Preconditioner(chol, stream), info

    return solve_with_preconditioner
