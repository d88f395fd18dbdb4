"""Matrix-exponential action through the Arnoldi factorisation and its gradient through the adjoint
sweep: the callers of `arnoldi.hessenberg` in the reference's PDE experiments
(`/root/reference/src/matfree_extensions/util/pde_util.py:240-268`), behind the same factory names.

    expm = pde.expm_arnoldi(10)
    solve = pde.solver_expm(0.0, 1.0, wave_operator, expm=expm)
    (y1, info), pullback = bl.vjp(solve, y0, scale)
    dy0, dscale = pullback(u)                    # cotangent u of y1

The n-sized work (the factorisation, `Q y`, `Q^T u`, the rank-one basis cotangent) runs on the
device; the K x K `expm` and its Fréchet derivative stay on the host (SciPy), like the K x K `eigh`
of the SLQ integrand.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg

from experiments_lanczos_adjoints_b200 import _lib, arnoldi
from experiments_lanczos_adjoints_b200 import device as dev
from experiments_lanczos_adjoints_b200.lanczos import _vec_ws
from experiments_lanczos_adjoints_b200.operators import WaveStencilOperator


def stencil_laplacian(dx):
    """`pde_util.stencil_laplacian` (`pde_util.py:18-20`)."""
    return WaveStencilOperator.stencil_laplacian(dx)


def pde_wave_anisotropic(scale_like, /, stencil, *, constrain="square", boundary="neumann"):
    """`pde_util.pde_wave_anisotropic` (`pde_util.py:126-143`) as an operator object: returns
    `(operator, {"scale": empty_like(scale_like)})`; `operator(y, scale)` is the reference's
    `parametrize(scale=scale)(y)` for flat `y = (u, du)`.  The kernel hard-wires what the reference's
    training script uses: `constrain = square` (`train.py:57-59`) and edge-replicating (Neumann)
    padding (`pde_util.py:153-157`)."""
    if constrain not in ("square", np.square) or boundary != "neumann":
        raise NotImplementedError("the stencil kernel implements constrain=square with the Neumann boundary")
    shape = np.shape(scale_like)
    if len(shape) != 2 or shape[0] != shape[1]:
        raise ValueError("scale must be a square (g, g) field")
    return WaveStencilOperator(shape[0], stencil), {"scale": np.empty(shape, dtype=np.asarray(scale_like).dtype)}


def _rows_dot(basis, x, stream):
    """`basis @ x` for a `(K, n)` basis: K device dot products, returned on the host in float64."""
    K, n = basis._shape
    out = dev.DeviceArray((K,), basis.dtype)
    ws, nbytes = _vec_ws()
    _lib.call("bl_rows_dot", dev.dtype_code(basis.dtype), n, K, basis.ptr, basis.ld, x.ptr, out.ptr, ws.ptr, nbytes,
              stream.ptr)  # fmt: skip
    return out.numpy(stream).astype(np.float64)


def _rows_combine(basis, coef, stream):
    """`basis.T @ coef`: a combination of basis rows with host coefficients."""
    K, n = basis._shape
    out = dev.DeviceArray((n,), basis.dtype)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    ws, nbytes = _vec_ws()
    _lib.call("bl_rows_combine", dev.dtype_code(basis.dtype), n, K, basis.ptr, basis.ld, coef.ctypes.data, 0, out.ptr,
              ws.ptr, nbytes, stream.ptr)  # fmt: skip
    stream.synchronize()  # `coef` is a host temporary
    return out


class _ExpmArnoldi:
    def __init__(self, krylov_depth, max_squarings, reortho, custom_vjp):
        self.K, self.max_squarings = krylov_depth, max_squarings
        self.kwargs = {"reortho": reortho, "custom_vjp": custom_vjp}

    def _forward(self, Qt, H, c, dt, stream):
        basis = Qt.T  # storage layout (K, n)
        Hh, ch = H.numpy(stream).astype(np.float64), float(c.numpy(stream))
        y = scipy.linalg.expm(dt * Hh)[:, 0]  # expm(dt H) e1                                pde_util.py:264-265
        return _rows_combine(basis, y / ch, stream), (basis, Hh, ch, y)  # 1/c * Q @ expmat @ e1  :266

    def __call__(self, matvec, dt, y0_flat, *p, stream=None):
        stream = stream or dev.default_stream()
        algorithm = arnoldi.hessenberg(matvec, self.K, **self.kwargs)
        Qt, H, _r, c = algorithm(y0_flat, *p, stream=stream)
        out, _ = self._forward(Qt, H, c, float(dt), stream)
        return out, {"num_matvecs": self.K}

    def vjp(self, matvec, dt, y0_flat, *p, stream=None):
        stream = stream or dev.default_stream()
        dt = float(dt)
        algorithm = arnoldi.hessenberg(matvec, self.K, **self.kwargs)
        (Qt, H, _r, c), pull = algorithm.vjp(y0_flat, *p, stream=stream)
        out, (basis, Hh, ch, y) = self._forward(Qt, H, c, dt, stream)
        K, n = basis._shape

        def pullback(cotangent):
            u = cotangent[0] if isinstance(cotangent, tuple) else cotangent  # (out, info): info carries none
            u = dev.asarray(u if isinstance(u, dev.DeviceArray) else np.asarray(u).reshape(-1), dtype=basis.dtype)
            qtu = _rows_dot(basis, u, stream)  # Q^T u
            # out = (1/c) Q y,  y = expm(dt H) e1
            dy = qtu / ch
            e1 = np.zeros(K)
            e1[0] = 1.0
            dH = dt * scipy.linalg.expm_frechet(dt * Hh.T, np.outer(dy, e1), compute_expm=False)
            dc = -float(qtu @ y) / ch**2
            dQ = dev.DeviceArray((K, n), basis.dtype, ld=basis.ld)  # rank one: row k = (y_k / c) u
            for k in range(K):
                _lib.call("bl_vec_axpby", dev.dtype_code(basis.dtype), n, float(y[k] / ch), u.ptr, 0.0, None,
                          dQ.row(k).ptr, stream.ptr)  # fmt: skip
            return pull((dQ.T, dH.astype(basis.dtype), None, np.asarray(dc, dtype=basis.dtype)))

        return (out, {"num_matvecs": self.K}), pullback


def expm_arnoldi(krylov_depth, *, max_squarings: int = 32, reortho="full", custom_vjp=True):
    """Drop-in for `pde_util.expm_arnoldi` (`pde_util.py:257-268`): `expm(matvec, dt, y0_flat, *p)`
    returns `(1/c * Q @ expm(dt * H) @ e1, {"num_matvecs": krylov_depth})`.  `max_squarings` is
    accepted for signature compatibility (SciPy's scaling-and-squaring picks its own)."""
    return _ExpmArnoldi(krylov_depth, max_squarings, reortho, custom_vjp)


class _SolverExpm:
    def __init__(self, t0, t1, vector_field, expm):
        arnoldi._require_operator(vector_field)
        self.dt, self.field, self.expm = t1 - t0, vector_field, expm

    @staticmethod
    def _flat(y0):
        if isinstance(y0, dev.DeviceArray):
            return y0, (y0.shape if y0.ndim == 1 else None)
        y0 = np.asarray(y0)
        return y0.reshape(-1), y0.shape

    def __call__(self, y0, *p):
        flat, _ = self._flat(y0)  # pde_util.py:244 ravel_pytree; the result stays flat on the device
        return self.expm(self.field, self.dt, flat, *p)

    def vjp(self, y0, *p):
        flat, _ = self._flat(y0)
        return self.expm.vjp(self.field, self.dt, flat, *p)


class _BatchedSolverExpm:
    """`jax.vmap(solve, in_axes=(0, None))(y0s, *p)` of the reference's training loss
    (`/root/reference/experiments/applications/partial_differential_equation/train.py:104-110`): a stack of initial
    conditions sharing one parameter set.  The B Arnoldi runs advance in lockstep
    (`bl_arnoldi_forward_batch` / `bl_arnoldi_adjoint_batch`); the parameter cotangent returned by the pullback is
    the SUM over the batch (what autodiff of a scalar loss of the stacked outputs gives), `dy0s` is per run."""

    def __init__(self, solver):
        if not isinstance(solver.expm, _ExpmArnoldi):
            raise TypeError("vmap(solve) is implemented for expm_arnoldi")
        self.solver = solver

    def _forward(self, y0s, p, stream):
        expm, op, dt = self.solver.expm, self.solver.field, float(self.solver.dt)
        y0s = np.asarray(y0s)
        B = y0s.shape[0]
        flat = np.ascontiguousarray(y0s.reshape(B, -1))
        n, dtype, K = flat.shape[1], flat.dtype, expm.K
        if not isinstance(K, (int, np.integer)) or K < 1 or K > n:
            raise ValueError(f"Parameter depth {K} is outside the expected range")
        if n != op.n:
            raise ValueError(f"operator acts on vectors of length {op.n}, got {n}")
        alg = arnoldi.hessenberg(op, K, **expm.kwargs)  # validates reortho like the reference (arnoldi.py:16-19)
        bound = op.bind(p, dtype, stream)
        code, ld = dev.dtype_code(dtype), dev.basis_ld(n, dtype)
        per = _lib.load().bl_arnoldi_workspace_bytes(n, K, code)
        V = dev.asarray(flat)
        Q = dev.DeviceArray((B * K, n), dtype, ld=ld)
        H = dev.DeviceArray((B, K * K), dtype)
        r = dev.DeviceArray((B, n), dtype, ld=ld)
        c = dev.DeviceArray((B,), dtype)
        ws = dev.DeviceArray(((per * B + 7) // 8,), np.float64)
        arnoldi._run(op, "bl_arnoldi_forward_batch", op._handle, code, n, K, alg._forward_flags, B, V.ptr, n, Q.ptr, ld,
                     H.ptr, r.ptr, c.ptr, ws.ptr, per * B, stream.ptr)  # fmt: skip
        Hh = H.numpy(stream).reshape(B, K, K).astype(np.float64)
        ch = c.numpy(stream).astype(np.float64)
        out = dev.DeviceArray((B, n), dtype, ld=ld)
        ys = np.zeros((B, K))
        vws, vbytes = _vec_ws()
        for b in range(B):
            ys[b] = scipy.linalg.expm(dt * Hh[b])[:, 0]  # expm(dt H) e1                    pde_util.py:264-265
            coef = np.ascontiguousarray(ys[b] / ch[b], dtype=np.float64)  # 1/c * Q @ expmat @ e1      :266
            _lib.call("bl_rows_combine", code, n, K, Q.ptr + b * K * ld * dtype.itemsize, ld, coef.ctypes.data, 0,
                      out.row(b).ptr, vws.ptr, vbytes, stream.ptr)  # fmt: skip
            stream.synchronize()  # `coef` is a host temporary
        return out, (alg, bound, B, n, K, dtype, code, ld, per, Q, H, r, c, ws, Hh, ch, ys)

    def __call__(self, y0s, *p, stream=None):
        stream = stream or dev.default_stream()
        out, _ = self._forward(y0s, p, stream)
        return out, {"num_matvecs": self.solver.expm.K}

    def vjp(self, y0s, *p, stream=None):
        stream = stream or dev.default_stream()
        op, dt = self.solver.field, float(self.solver.dt)
        out, (alg, bound, B, n, K, dtype, code, ld, per, Q, H, r, c, ws, Hh, ch, ys) = self._forward(y0s, p, stream)
        if not alg.custom_vjp:
            raise NotImplementedError("autodiff through the loop is not available; use custom_vjp=True")
        item = dtype.itemsize

        def pullback(cotangent):
            u = cotangent[0] if isinstance(cotangent, tuple) else cotangent  # (outs, info): info carries none
            U = dev.asarray(np.ascontiguousarray(np.asarray(u, dtype=dtype).reshape(B, n)))
            dH = np.zeros((B, K, K), dtype)
            dc = np.zeros(B, dtype)
            dQ = dev.DeviceArray((B * K, n), dtype, ld=ld)  # rank one per run: row k = (y_k / c) u
            qtu_d = dev.DeviceArray((K,), dtype)
            vws, vbytes = _vec_ws()
            e1 = np.zeros(K)
            e1[0] = 1.0
            for b in range(B):
                ub = U.row(b)
                _lib.call("bl_rows_dot", code, n, K, Q.ptr + b * K * ld * item, ld, ub.ptr, qtu_d.ptr, vws.ptr, vbytes,
                          stream.ptr)  # fmt: skip
                qtu = qtu_d.numpy(stream).astype(np.float64)  # Q^T u
                dH[b] = dt * scipy.linalg.expm_frechet(dt * Hh[b].T, np.outer(qtu / ch[b], e1), compute_expm=False)
                dc[b] = -float(qtu @ ys[b]) / ch[b] ** 2
                for k in range(K):
                    _lib.call("bl_vec_axpby", code, n, float(ys[b, k] / ch[b]), ub.ptr, 0.0, None,
                              dQ.ptr + (b * K + k) * ld * item, stream.ptr)  # fmt: skip
            dHd, dcd = dev.asarray(dH.reshape(B, K * K)), dev.asarray(dc)
            op.bind(bound, dtype, stream)  # same parameter values as the forward pass
            op.grad_zero(dtype, stream)
            dv = dev.DeviceArray((B, n), dtype, ld=ld)
            Lam = dev.DeviceArray((B * K, n), dtype, ld=ld)
            arnoldi._run(op, "bl_arnoldi_adjoint_batch", op._handle, code, n, K, alg._adjoint_flags, B, Q.ptr, ld, H.ptr,
                         r.ptr, c.ptr, dQ.ptr, dHd.ptr, None, dcd.ptr, dv.ptr, ld, Lam.ptr, ws.ptr, per * B,
                         stream.ptr)  # fmt: skip
            grads = op.grad_export(dtype, stream=stream)
            return (dv, *grads)

        return (out, {"num_matvecs": K}), pullback


def vmap_solver(solve):
    """`jax.vmap(solve, in_axes=(0, None))` for a `solver_expm` object: `vmap_solver(solve)(y0s, *p)` returns the
    stacked outputs `(B, n)`; `bl.vjp(vmap_solver(solve), y0s, *p)` the pullback `(dy0s (B, n), *dparams)`."""
    return _BatchedSolverExpm(solve)


def solver_expm(t0, t1, vector_field, /, expm):
    """Drop-in for `pde_util.solver_expm` (`pde_util.py:240-254`) with an operator object as the
    vector field: `solve(y0, *p) -> (y1_flat, info)`; `bl.vjp(solve, y0, *p)` gives the pullback."""
    return _SolverExpm(t0, t1, vector_field, expm)
