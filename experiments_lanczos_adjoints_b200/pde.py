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


def solver_expm(t0, t1, vector_field, /, expm):
    """Drop-in for `pde_util.solver_expm` (`pde_util.py:240-254`) with an operator object as the
    vector field: `solve(y0, *p) -> (y1_flat, info)`; `bl.vjp(solve, y0, *p)` gives the pullback."""
    return _SolverExpm(t0, t1, vector_field, expm)
