"""Pre-planned forward + adjoint of `lanczos.tridiag(..., reortho="full")`.

The function objects in `arnoldi.py` / `lanczos.py` mirror the reference's call structure
(allocate outputs per call, return arrays).  A training / SLQ loop calls the same forward and
adjoint thousands of times on one operand shape; this plan owns every buffer (basis `Q`,
adjoint basis `Lambda`, `H`, workspace, gradient) once and re-enqueues the two C-ABI calls —
no allocation, no host synchronisation inside `run()`.  Same arithmetic, same C entry points
(`bl_arnoldi_forward`, `bl_arnoldi_adjoint`).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from experiments_lanczos_adjoints_b200 import _lib
from experiments_lanczos_adjoints_b200 import device as dev
from experiments_lanczos_adjoints_b200.arnoldi import adjoint_flags, forward_flags


def pinned_empty(shape, dtype) -> np.ndarray:
    """NumPy array backed by page-locked host memory (fast, truly asynchronous H2D/D2H)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = C.c_void_p()
    _lib.call("bl_host_alloc", C.byref(p), max(1, nbytes))
    buf = (C.c_char * max(1, nbytes)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    _PINNED[arr.ctypes.data] = p.value  # freed at interpreter exit with the CUDA context
    return arr


_PINNED = {}


def pinned_free(arr: np.ndarray) -> None:
    """Give back a buffer of `pinned_empty` (the caller guarantees that no copy from or to it is in flight)."""
    p = _PINNED.pop(arr.ctypes.data, None)
    if p is not None:
        _lib.call("bl_host_free", p)


class PinnedBuffers:
    """Pinned arrays that live as long as their owner (a cache entry of an estimator): freed with it."""

    def __init__(self, shapes, dtype):
        self.arrays = [pinned_empty(shape, dtype) for shape in shapes]

    def __getitem__(self, i):
        return self.arrays[i]

    def __del__(self):
        try:
            dev.synchronize()  # copies out of the buffers may still be enqueued
            for a in self.arrays:
                pinned_free(a)
        except Exception:  # interpreter shutdown: the CUDA context takes the memory with it
            pass
        self.arrays = []


class TridiagAdjointPlan:
    """`(Q^T, alpha, beta), r = tridiag(op, K, reortho="full")(v, params)` followed by the adjoint
    for cotangents on `(alpha, beta)` (the SLQ case) or on every output."""

    def __init__(self, op, krylov_depth: int, dtype, stream: dev.Stream | None = None, tridiagonal_cotangent: bool = True):
        self.op, self.K, self.dtype = op, int(krylov_depth), np.dtype(dtype)
        self.n = op.n
        self.stream = stream or dev.default_stream()
        n, K = self.n, self.K
        if K < 1 or K > n:
            raise ValueError(f"Parameter depth {K} is outside the expected range")
        self.code = dev.dtype_code(self.dtype)
        # `tridiag` is defined for symmetric operands and its cotangent dH is tridiagonal: the adjoint treats H as
        # tridiagonal and Gamma as banded (BL_ADJ_SYMMETRIC | BL_ADJ_TRIDIAG_COTANGENT, include/b200_lanczos.h)
        self.adjoint_flags = adjoint_flags(True, True, tridiagonal_cotangent)
        self._symmetric_flags = self.adjoint_flags
        self.ld = dev.basis_ld(n, self.dtype)
        self.Q = dev.DeviceArray((K, n), self.dtype, ld=self.ld)
        self.Lam = dev.DeviceArray((K, n), self.dtype, ld=self.ld)
        self.H = dev.DeviceArray((K, K), self.dtype)
        self.dH = dev.DeviceArray((K, K), self.dtype)
        self.r = dev.DeviceArray((n,), self.dtype)
        self.c = dev.DeviceArray((), self.dtype)
        self.v = dev.DeviceArray((n,), self.dtype)
        self.dv = dev.DeviceArray((n,), self.dtype)
        self.params = [dev.DeviceArray(tuple(s) or (1,), self.dtype) for s in op.param_shapes()]
        self.grads = [dev.DeviceArray(tuple(s) or (1,), self.dtype) for s in op.param_shapes()]
        self.ws_bytes = _lib.load().bl_arnoldi_workspace_bytes(n, K, self.code)
        self.ws = dev.DeviceArray(((self.ws_bytes + 3) // 4,), np.float32)
        self._pptr = (C.c_void_p * max(1, len(self.params)))(*[p.ptr for p in self.params])
        self._gptr = (C.c_void_p * max(1, len(self.grads)))(*[g.ptr for g in self.grads])

    # -- uploads (async on the plan's stream; pass pinned arrays for true overlap) ----------
    def _h2d(self, dst: dev.DeviceArray, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        _lib.call("bl_memcpy_h2d", dst.ptr, host.ctypes.data, host.nbytes, self.stream.ptr)
        return host.nbytes

    def set_vector(self, v_host):
        return self._h2d(self.v, v_host)

    def set_params(self, *params_host):
        return sum(self._h2d(d, h) for d, h in zip(self.params, params_host))

    def set_cotangent(self, dH_host):
        return self._h2d(self.dH, dH_host)

    # -- the two sweeps -----------------------------------------------------------------------
    def forward(self):
        s = self.stream.ptr
        _lib.call("bl_op_set_params", self.op._handle, self.code, self._pptr, len(self.params), s)
        _lib.call("bl_arnoldi_forward", self.op._handle, self.code, self.n, self.K, forward_flags(True, True), self.v.ptr, self.Q.ptr,
                  self.ld, self.H.ptr, self.r.ptr, self.c.ptr, self.ws.ptr, self.ws_bytes, s)  # fmt: skip

    def adjoint(self, dQ=None, dr=None, zero=True, export=True):
        """`zero=False` keeps accumulating the parameter cotangent inside the operator (a sum over probes),
        `export=False` leaves it there; `export_grads()` brings the sum out once."""
        s = self.stream.ptr
        if zero:
            _lib.call("bl_op_grad_zero", self.op._handle, self.code, s)
        _lib.call("bl_arnoldi_adjoint", self.op._handle, self.code, self.n, self.K, self.adjoint_flags, self.Q.ptr, self.ld,
                  self.H.ptr, self.r.ptr, self.c.ptr, dQ.ptr if dQ is not None else None, self.dH.ptr,
                  dr.ptr if dr is not None else None, None, self.dv.ptr, self.Lam.ptr, self.ws.ptr,
                  self.ws_bytes, s)  # fmt: skip
        if export:
            self.export_grads()

    def export_grads(self):
        _lib.call("bl_op_grad_export", self.op._handle, self.code, self._gptr, len(self.grads), self.stream.ptr)

    def run(self):
        """One forward + adjoint, enqueued back to back (no host sync)."""
        self.forward()
        self.adjoint()

    # -- host-buffer entry point (the e2e path of bench.py) -----------------------------------
    def run_host(self, v_host, params_host, dH_host, out_H, out_dv, out_grads, sync=True):
        """Host buffers in, host buffers out: H2D of `(v, params, dH)`, forward + adjoint, D2H of
        `(H, dv, dparams)`; synchronises once at the end (`sync=False`: the caller synchronises the plan's
        stream -- several plans on several streams overlap one probe's copies with another's kernels).
        Returns (h2d_bytes, d2h_bytes)."""
        h2d = self.set_vector(v_host) + self.set_params(*params_host) + self.set_cotangent(dH_host)
        self.run()
        d2h = 0
        for src, dst in [(self.H, out_H), (self.dv, out_dv), *zip(self.grads, out_grads)]:
            _lib.call("bl_memcpy_d2h", dst.ctypes.data, src.ptr, dst.nbytes, self.stream.ptr)
            d2h += dst.nbytes
        if sync:
            self.stream.synchronize()
        return h2d, d2h

    def coefficients(self):
        """`(alpha, beta)` of the last forward (`lanczos.py:162-164`); synchronises.  Also checks on this `H` that
        the operand behaved like a symmetric one (`lanczos.hessenberg_is_tridiagonal`): if not, the next `adjoint()`
        runs the general Arnoldi loops (what the reference computes for any operand).  `run()` -- forward and
        adjoint without a host read in between -- cannot check and is for operands known to be symmetric."""
        import warnings

        from experiments_lanczos_adjoints_b200.lanczos import hessenberg_is_tridiagonal

        H = self.H.numpy(self.stream)
        if hessenberg_is_tridiagonal(H):
            self.adjoint_flags = self._symmetric_flags
        else:
            warnings.warn("TridiagAdjointPlan: operand is not symmetric on this Krylov space; general adjoint loops",
                          stacklevel=2)  # fmt: skip
            self.adjoint_flags = adjoint_flags(True, False, False)
        T = 0.5 * (H + H.T)
        return np.diag(T, 0).copy(), np.diag(T, 1).copy()


class BatchedTridiagAdjointPlan:
    """`TridiagAdjointPlan` for P independent start vectors (Hutchinson probes) sharing one operand: the P runs
    advance in LOCKSTEP through `bl_arnoldi_forward_batch` / `bl_arnoldi_adjoint_batch` -- per Krylov step one
    batched operator call (a sparse operand's values and indices are read once for all runs) and one Gram-Schmidt
    step kernel for every four runs (`k_step_tma`: the grid-wide reductions of a step are shared by the runs).  This
    is the batched handler behind the reference's `jax.vmap(integrand)` (hutchinson.py:14,53).  The parameter
    cotangent that comes out is the SUM over the P runs (what the Hutchinson mean needs, hutchinson.py:54)."""

    def __init__(self, op, krylov_depth: int, dtype, count: int, stream: dev.Stream | None = None):
        self.op, self.K, self.dtype, self.P = op, int(krylov_depth), np.dtype(dtype), int(count)
        self.n = op.n
        self.stream = stream or dev.default_stream()
        n, K, P = self.n, self.K, self.P
        if K < 1 or K > n:
            raise ValueError(f"Parameter depth {K} is outside the expected range")
        if P < 1:
            raise ValueError("count must be positive")
        self.code = dev.dtype_code(self.dtype)
        self.forward_flags = forward_flags(True, True)
        self.adjoint_flags = adjoint_flags(True, True, True)
        self.ld = dev.basis_ld(n, self.dtype)
        self.Q = dev.DeviceArray((P * K, n), self.dtype, ld=self.ld)
        self.Lam = dev.DeviceArray((P * K, n), self.dtype, ld=self.ld)
        self.H = dev.DeviceArray((P, K * K), self.dtype)
        self.dH = dev.DeviceArray((P, K * K), self.dtype)
        self.r = dev.DeviceArray((P, n), self.dtype, ld=self.ld)
        self.c = dev.DeviceArray((P,), self.dtype)
        self.v = dev.DeviceArray((P, n), self.dtype)
        self.dv = dev.DeviceArray((P, n), self.dtype)
        self.params = [dev.DeviceArray(tuple(s) or (1,), self.dtype) for s in op.param_shapes()]
        self.grads = [dev.DeviceArray(tuple(s) or (1,), self.dtype) for s in op.param_shapes()]
        self.ws_bytes = P * _lib.load().bl_arnoldi_workspace_bytes(n, K, self.code)
        self.ws = dev.DeviceArray(((self.ws_bytes + 3) // 4,), np.float32)
        self._pptr = (C.c_void_p * max(1, len(self.params)))(*[p.ptr for p in self.params])
        self._gptr = (C.c_void_p * max(1, len(self.grads)))(*[g.ptr for g in self.grads])

    def _h2d(self, dst: dev.DeviceArray, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        _lib.call("bl_memcpy_h2d", dst.ptr, host.ctypes.data, host.nbytes, self.stream.ptr)
        return host.nbytes

    def set_vectors(self, v_host):  # (P, n)
        return self._h2d(self.v, v_host)

    def set_params(self, *params_host):
        return sum(self._h2d(d, h) for d, h in zip(self.params, params_host))

    def set_cotangents(self, dH_host):  # (P, K, K)
        return self._h2d(self.dH, dH_host)

    def forward(self):
        s = self.stream.ptr
        _lib.call("bl_op_set_params", self.op._handle, self.code, self._pptr, len(self.params), s)
        _lib.call("bl_arnoldi_forward_batch", self.op._handle, self.code, self.n, self.K, self.forward_flags, self.P, self.v.ptr,
                  self.n, self.Q.ptr, self.ld, self.H.ptr, self.r.ptr, self.c.ptr, self.ws.ptr, self.ws_bytes, s)  # fmt: skip

    def adjoint(self, zero=True, export=True, general=False):
        """`general=True`: the general Arnoldi adjoint loops (an operand that is not symmetric, see `coefficients`)."""
        s = self.stream.ptr
        if zero:
            _lib.call("bl_op_grad_zero", self.op._handle, self.code, s)
        flags = adjoint_flags(True, False, False) if general else self.adjoint_flags
        _lib.call("bl_arnoldi_adjoint_batch", self.op._handle, self.code, self.n, self.K, flags, self.P, self.Q.ptr, self.ld,
                  self.H.ptr, self.r.ptr, self.c.ptr, None, self.dH.ptr, None, None, self.dv.ptr, self.n, self.Lam.ptr,
                  self.ws.ptr, self.ws_bytes, s)  # fmt: skip
        if export:
            self.export_grads()

    def export_grads(self):
        _lib.call("bl_op_grad_export", self.op._handle, self.code, self._gptr, len(self.grads), self.stream.ptr)

    def run(self):
        """Forward + adjoint of all P runs, enqueued back to back (no host sync)."""
        self.forward()
        self.adjoint()

    def run_host(self, v_host, params_host, dH_host, out_H, out_dv, out_grads, sync=True, adjoint_first=False):
        """Host buffers in, host buffers out (see `TridiagAdjointPlan.run_host`); `v_host (P, n)`, `dH_host (P, K, K)`,
        `out_H (P, K, K)`, `out_dv (P, n)`.  Returns (h2d_bytes, d2h_bytes).

        `adjoint_first`: the software-pipelined order of a lane that runs half a cycle behind its neighbour (see
        `bench.py`): adjoint of the PREVIOUS call's forward, its results to the host, then this call's inputs to the device
        and their forward.  Same copies, same kernels per call; the outputs are those of the previous call's inputs."""
        d2h = 0

        def results():
            nonlocal d2h
            for src, dst in [(self.H, out_H), (self.dv, out_dv), *zip(self.grads, out_grads)]:
                _lib.call("bl_memcpy_d2h", dst.ctypes.data, src.ptr, dst.nbytes, self.stream.ptr)
                d2h += dst.nbytes

        if adjoint_first:
            self.adjoint()
            results()
            h2d = self.set_vectors(v_host) + self.set_params(*params_host) + self.set_cotangents(dH_host)
            self.forward()
        else:
            h2d = self.set_vectors(v_host) + self.set_params(*params_host) + self.set_cotangents(dH_host)
            self.run()
            results()
        if sync:
            self.stream.synchronize()
        return h2d, d2h

    def coefficients(self):
        """`[(alpha, beta)] * P` of the last forward (`lanczos.py:162-164`) and whether every `H` is tridiagonal up to
        rounding (`lanczos.hessenberg_is_tridiagonal`: the operand behaved like a symmetric one); synchronises."""
        from experiments_lanczos_adjoints_b200.lanczos import hessenberg_is_tridiagonal

        H = self.H.numpy(self.stream).reshape(self.P, self.K, self.K)
        out, symmetric = [], True
        for Hp in H:
            T = 0.5 * (Hp + Hp.T)
            out.append((np.diag(T, 0).copy(), np.diag(T, 1).copy()))
            symmetric = symmetric and hessenberg_is_tridiagonal(Hp)
        return out, symmetric


def profile(fn):
    """Run `fn()` with per-kernel-class event timing; returns
    `{class: {"launches", "ms", "algorithmic_bytes"}}` (see `bl_profile_begin` in the header)."""
    names = ["dots", "combine", "matvec", "vjp", "other", "fused"]
    _lib.call("bl_profile_begin")
    try:
        fn()
    finally:
        counts = (C.c_uint64 * len(names))()
        ms = (C.c_double * len(names))()
        nbytes = (C.c_double * len(names))()
        _lib.call("bl_profile_end", counts, ms, nbytes)
    return {nm: {"launches": int(counts[i]), "ms": float(ms[i]), "algorithmic_bytes": float(nbytes[i])}
            for i, nm in enumerate(names)}  # fmt: skip
