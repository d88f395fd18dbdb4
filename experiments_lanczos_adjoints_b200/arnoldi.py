"""Arnoldi / Hessenberg factorisation with CGS2 re-orthogonalisation and its adjoint.

Host-side mirror of `/root/reference/src/matfree_extensions/arnoldi.py`: same factory
signature, same outputs `(Q (n,K), H (K,K), r (n,), c ())`, same error behaviour.  The two
loops (`_forward`, `arnoldi.py:57-101`; `_adjoint`, `arnoldi.py:104-220`) run in
`bl_arnoldi_forward` / `bl_arnoldi_adjoint` of libb200lanczos.so.
"""

from __future__ import annotations

import numpy as np

from experiments_lanczos_adjoints_b200 import _lib
from experiments_lanczos_adjoints_b200 import device as dev
from experiments_lanczos_adjoints_b200.operators import Operator


def _require_operator(matvec):
    if not isinstance(matvec, Operator):
        raise TypeError(
            "matvec must be an operator object from experiments_lanczos_adjoints_b200.operators "
            "(SparseOperator, DenseOperator, GramOperator, WaveStencilOperator, CallbackOperator); "
            "plain Python callables cannot be traced onto the device without JAX."
        )


def _run(op, name, *args):
    """C-ABI call that re-raises an exception thrown inside a user callback."""
    try:
        _lib.call(name, *args)
    except _lib.BLError:
        err, op._error = getattr(op, "_error", None), None
        if err is not None:
            raise err
        raise


BL_ADJ_REORTHO_FULL, BL_ADJ_SYMMETRIC, BL_ADJ_TRIDIAG_COTANGENT = 1, 2, 4  # include/b200_lanczos.h


def forward_flags(second_pass: bool, symmetric: bool) -> int:
    """The `second_pass` argument of `bl_arnoldi_forward(_batch)`."""
    return (1 if second_pass else 0) | (2 if second_pass and symmetric else 0)


def adjoint_flags(reortho_full: bool, symmetric: bool, tridiagonal_cotangent: bool = False) -> int:
    """The `reortho_full` argument of `bl_arnoldi_adjoint(_batch)`."""
    flags = BL_ADJ_REORTHO_FULL if reortho_full else 0
    if reortho_full and symmetric:
        flags |= BL_ADJ_SYMMETRIC
        if tridiagonal_cotangent:
            flags |= BL_ADJ_TRIDIAG_COTANGENT
    return flags


def _ptr(a):
    return a.ptr if a is not None else None


class _Workspace:
    """Per-(n, K, dtype) scratch, reused across calls of one algorithm object."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, nbytes):
        buf = self._bufs.get(key)
        if buf is None or buf.size * 4 < nbytes:
            buf = dev.DeviceArray(((nbytes + 3) // 4,), np.float32)  # raw bytes
            self._bufs[key] = buf
        return buf


def _cotangent_basis(x, K, n, dtype, transposed_input):
    """Bring a `(n, K)` (reference layout) or `(K, n)` cotangent into basis layout."""
    if x is None:
        return None
    if isinstance(x, dev.DeviceArray):
        if x.ndim != 2:
            raise ValueError("basis cotangent must be 2-D")
        stored_kn = x._shape == (K, n)
        if stored_kn and x.dtype == np.dtype(dtype) and (x.ld * x.dtype.itemsize) % 16 == 0 and (
            x.is_transposed == transposed_input
        ):
            return x  # already K rows of length n in device memory
        x = x.numpy()
    mat = np.asarray(x, dtype=dtype)
    if transposed_input:
        mat = mat.T  # (n, K) -> (K, n)
    if mat.shape != (K, n):
        raise ValueError(f"basis cotangent has shape {np.shape(x)}, expected {(n, K) if transposed_input else (K, n)}")
    if not mat.any():
        return None
    return dev.basis_from_host(mat, dtype)


def _cotangent_vec(x, n, dtype):
    if x is None:
        return None
    if isinstance(x, dev.DeviceArray):
        return dev.asarray(x, dtype=dtype)
    arr = np.asarray(x, dtype=dtype).reshape(-1)
    if arr.size != n:
        raise ValueError(f"cotangent has {arr.size} elements, expected {n}")
    if not arr.any():
        return None
    return dev.asarray(arr)


class HessenbergEstimate:
    """What `arnoldi.hessenberg(...)` returns: `estimate(v, *params) -> (Q, H, r, c)`,
    plus `.vjp(v, *params) -> (outputs, pullback)` standing in for `jax.vjp` on the
    reference's `custom_vjp` pair (`arnoldi.py:29-53`)."""

    def __init__(self, op, krylov_depth, *, reortho, custom_vjp, reortho_vjp):
        self.op, self.K = op, krylov_depth
        self.reortho, self.custom_vjp, self.reortho_vjp = reortho, custom_vjp, reortho_vjp
        # set by `lanczos.tridiag(reortho="full")`, whose operand is symmetric by contract: the adjoint may
        # treat H as tridiagonal (BL_ADJ_SYMMETRIC in include/b200_lanczos.h)
        self.symmetric = False
        self.tridiagonal_cotangent = False  # dH handed to the pullback is tridiagonal (BL_ADJ_TRIDIAG_COTANGENT)
        self._ws = _Workspace()

    # arnoldi.py:26 — `reortho_` is always `reortho_vjp`; only "none" switches the 2nd pass off
    @property
    def _second_pass(self) -> bool:
        return self.reortho_vjp != "none"

    @property
    def _forward_flags(self) -> int:  # BL_FWD_SECOND_PASS | BL_FWD_SYMMETRIC (include/b200_lanczos.h)
        return forward_flags(self._second_pass, self.symmetric)

    @property
    def _adjoint_flags(self) -> int:
        return adjoint_flags(self.reortho == "full", self.symmetric and self._second_pass, self.tridiagonal_cotangent)

    def _forward(self, v, params, stream):
        op, K = self.op, self.K
        v = dev.asarray(v)
        if v.ndim != 1:
            raise ValueError("v must be a flat vector")
        n, dtype = v.shape[0], v.dtype
        if not isinstance(K, (int, np.integer)) or K < 1 or K > n:  # arnoldi.py:58-60
            raise ValueError(f"Parameter depth {K} is outside the expected range")
        if n != op.n:
            raise ValueError(f"operator acts on vectors of length {op.n}, got {n}")
        bound = op.bind(params, dtype, stream)
        ld = dev.basis_ld(n, dtype)
        Q = dev.DeviceArray((K, n), dtype, ld=ld)
        H = dev.DeviceArray((K, K), dtype)
        r = dev.DeviceArray((n,), dtype)
        c = dev.DeviceArray((), dtype)
        nbytes = _lib.load().bl_arnoldi_workspace_bytes(n, K, dev.dtype_code(dtype))
        ws = self._ws.get(("arnoldi", n, K, dtype.str), nbytes)
        _run(op, "bl_arnoldi_forward", op._handle, dev.dtype_code(dtype), n, K, self._forward_flags,
             v.ptr, Q.ptr, ld, H.ptr, r.ptr, c.ptr, ws.ptr, nbytes, stream.ptr)  # fmt: skip
        return (Q, H, r, c), (n, dtype, ld, nbytes, ws, bound)

    def __call__(self, v, *params, stream=None):
        if np.iscomplexobj(v) or any(np.iscomplexobj(p) for p in params if not isinstance(p, dev.DeviceArray)):
            return _forward_complex(self, v, params, stream or dev.default_stream())
        (Q, H, r, c), _ = self._forward(v, params, stream or dev.default_stream())
        return Q.T, H, r, c  # Q shown as (n, K) like the reference

    def vjp(self, v, *params, stream=None):
        if np.iscomplexobj(v) or any(np.iscomplexobj(p) for p in params if not isinstance(p, dev.DeviceArray)):
            raise NotImplementedError("complex inputs are supported in the forward only: the reference's adjoint "
                                      "transposes without conjugation (arnoldi.py:130,202-204)")  # fmt: skip
        if not self.custom_vjp:
            raise NotImplementedError(
                "custom_vjp=False asks for autodiff through the loop (arnoldi.py:51-53); this build "
                "has no tracing autodiff — differentiate with custom_vjp=True (the adjoint sweep)."
            )
        stream = stream or dev.default_stream()
        (Q, H, r, c), (n, dtype, ld, nbytes, ws, bound) = self._forward(v, params, stream)
        op, K = self.op, self.K

        def pullback(cotangents):
            dQ, dH, dr, dc = cotangents
            dQb = _cotangent_basis(dQ, K, n, dtype, transposed_input=True)
            dHd = dev.asarray(np.zeros((K, K), dtype) if dH is None else dH, dtype=dtype)
            drd = _cotangent_vec(dr, n, dtype)
            dcd = None if dc is None else dev.asarray(np.asarray(dc, dtype=dtype).reshape(1))
            op.bind(bound, dtype, stream)  # same parameter values as the forward pass
            op.grad_zero(dtype, stream)
            dv = dev.DeviceArray((n,), dtype)
            Lam = dev.DeviceArray((K, n), dtype, ld=ld)
            _run(op, "bl_arnoldi_adjoint", op._handle, dev.dtype_code(dtype), n, K,
                 self._adjoint_flags, Q.ptr, ld, H.ptr, r.ptr, c.ptr, _ptr(dQb), dHd.ptr,
                 _ptr(drd), _ptr(dcd), dv.ptr, Lam.ptr, ws.ptr, nbytes, stream.ptr)  # fmt: skip
            grads = op.grad_export(dtype, stream=stream)
            return (dv, *grads)

        return (Q.T, H, r, c), pullback


class ComplexDeviceArray:
    """A complex array held as two real device arrays (the library's kernels are real)."""

    def __init__(self, re: dev.DeviceArray, im: dev.DeviceArray):
        self.re, self.im = re, im
        self.shape = re.shape
        self.dtype = np.result_type(re.dtype, np.complex64)

    @property
    def T(self):
        return ComplexDeviceArray(self.re.T, self.im.T)

    def numpy(self, stream=None):
        return self.re.numpy(stream) + 1j * self.im.numpy(stream)


def _forward_complex(alg: "HessenbergEstimate", v, params, stream):
    """`arnoldi._forward` for a COMPLEX start vector / operand (`arnoldi.py:57-101` with its `.conj()`s, `:66,87,92,95`;
    the reference tests the forward with `dtype=complex`, `tests/test_arnoldi/test_hessenberg_forward.py:13`).  The
    library's kernels are real, so the step is composed on the host from real device calls on the real and imaginary
    parts -- four real matvecs per complex matvec, four blocks of row dots per `Q^H v`, four row combinations per
    `Q h` -- through the same C ABI (`bl_op_matvec`, `bl_rows_dot`, `bl_rows_combine`, `bl_vec_axpby`).  Forward only,
    one parameter array, sized for the reference's unit tests rather than for throughput."""
    from experiments_lanczos_adjoints_b200.lanczos import device_axpby, device_dot
    from experiments_lanczos_adjoints_b200.pde import _rows_combine, _rows_dot

    op, K = alg.op, alg.K
    v = np.asarray(v)
    if v.ndim != 1:
        raise ValueError("v must be a flat vector")
    n = v.shape[0]
    if not isinstance(K, (int, np.integer)) or K < 1 or K > n:  # arnoldi.py:58-60
        raise ValueError(f"Parameter depth {K} is outside the expected range")
    if len(params) != 1:
        raise NotImplementedError("the complex forward takes one parameter array")
    ctype = np.result_type(v.dtype, np.asarray(params[0]).dtype, np.complex64)
    rtype = np.float32 if ctype == np.complex64 else np.float64
    p = np.asarray(params[0], dtype=ctype)
    p_re, p_im = np.ascontiguousarray(p.real, dtype=rtype), np.ascontiguousarray(p.imag, dtype=rtype)
    v = v.astype(ctype)
    vr, vi = dev.asarray(np.ascontiguousarray(v.real, dtype=rtype)), dev.asarray(np.ascontiguousarray(v.imag, dtype=rtype))
    ld = dev.basis_ld(n, rtype)
    Qr, Qi = dev.zeros((K, n), rtype, ld=ld, stream=stream), dev.zeros((K, n), rtype, ld=ld, stream=stream)
    H = np.zeros((K, K), dtype=ctype)

    def matvec(xr, xi):  # (Ar + i Ai)(xr + i xi)
        a, b = op(xr, p_re), op(xi, p_im)
        c, d = op(xi, p_re), op(xr, p_im)
        return device_axpby(1.0, a, -1.0, b, stream), device_axpby(1.0, c, 1.0, d, stream)

    def project(xr, xi):  # Q^H x: <q_j, x> with the conjugate on q (arnoldi.py:87)
        return (_rows_dot(Qr, xr, stream) + _rows_dot(Qi, xi, stream)) + 1j * (_rows_dot(Qr, xi, stream) - _rows_dot(Qi, xr, stream))

    def subtract(xr, xi, h):  # x - Q h
        dr = device_axpby(1.0, _rows_combine(Qr, h.real, stream), -1.0, _rows_combine(Qi, h.imag, stream), stream)
        di = device_axpby(1.0, _rows_combine(Qr, h.imag, stream), 1.0, _rows_combine(Qi, h.real, stream), stream)
        return device_axpby(1.0, xr, -1.0, dr, stream), device_axpby(1.0, xi, -1.0, di, stream)

    def norm(xr, xi):  # sqrt(v^H v), arnoldi.py:66,95
        return float(np.sqrt(rtype(device_dot(xr, xr, stream) + device_dot(xi, xi, stream))))

    length0 = length = norm(vr, vi)
    for i in range(K):
        vr = device_axpby(1.0 / length, vr, 0.0, None, stream)  # arnoldi.py:80
        vi = device_axpby(1.0 / length, vi, 0.0, None, stream)
        item = np.dtype(rtype).itemsize
        for basis, part in ((Qr, vr), (Qi, vi)):  # Q[:, i] = v                   arnoldi.py:81
            _lib.call("bl_memcpy_d2d", basis.ptr + i * ld * item, part.ptr, n * item, stream.ptr)
        vr, vi = matvec(vr, vi)  # :84
        h = project(vr, vi).astype(ctype)  # :87 (rows > i of Q are zero)
        vr, vi = subtract(vr, vi, h)  # :88
        if alg._second_pass:  # :91-92 -- h is NOT updated
            vr, vi = subtract(vr, vi, project(vr, vi))
        length = norm(vr, vi)  # :95
        if i + 1 < K:  # :98 (the write at i+1 == K is dropped)
            h[i + 1] = length
        H[:, i] = h  # :99
    Q = ComplexDeviceArray(Qr, Qi).T  # shown as (n, K) like the reference
    return Q, H, ComplexDeviceArray(vr, vi), np.dtype(ctype).type(1.0 / length0)


def hessenberg(matvec, krylov_depth, /, *, reortho: str, custom_vjp: bool = True, reortho_vjp: str = "match"):
    """Drop-in for `arnoldi.hessenberg` (`/root/reference/src/matfree_extensions/arnoldi.py:7-54`)."""
    reortho_expected = ["none", "full"]
    if not isinstance(reortho, str) or reortho not in reortho_expected:  # arnoldi.py:16-19
        msg = f"Unexpected input for {reortho}: either of {reortho_expected} expected."
        raise TypeError(msg)
    _require_operator(matvec)
    return HessenbergEstimate(matvec, krylov_depth, reortho=reortho, custom_vjp=custom_vjp, reortho_vjp=reortho_vjp)
