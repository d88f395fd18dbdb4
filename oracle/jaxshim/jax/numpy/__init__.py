"""`jax.numpy` subset on torch tensors — see `jax/_core.py`."""

import numpy as _np
import torch

from jax.numpy import linalg  # noqa: F401

pi = _np.pi
float32 = torch.float32
float64 = torch.float64
int32 = torch.int32


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in x):
        return torch.stack([_t(e, dtype) for e in x])
    if isinstance(x, _np.ndarray) and x.dtype.kind == "f" and dtype is None:
        return torch.as_tensor(x)
    out = torch.as_tensor(x, dtype=dtype)
    if dtype is None and out.dtype == torch.float64 and not isinstance(x, _np.ndarray):
        out = out.to(torch.get_default_dtype())
    return out


def _dtype(dtype):
    if dtype is None:
        return torch.get_default_dtype()
    if dtype is float:
        return torch.get_default_dtype()
    if dtype is int:
        return torch.int64
    if dtype is complex:
        return torch.complex128 if torch.get_default_dtype() == torch.float64 else torch.complex64
    return dtype


def asarray(x, dtype=None):
    return _t(x, None if dtype is None else _dtype(dtype))


array = asarray


def shape(x):
    return tuple(_t(x).shape)


def ndim(x):
    return _t(x).ndim


def zeros(shape, dtype=None):
    return torch.zeros(shape, dtype=_dtype(dtype))


def ones(shape, dtype=None):
    return torch.ones(shape, dtype=_dtype(dtype))


def empty(shape, dtype=None):
    return torch.zeros(shape, dtype=_dtype(dtype))


def zeros_like(x):
    return torch.zeros_like(_t(x))


def ones_like(x):
    return torch.ones_like(_t(x))


def empty_like(x):
    return torch.zeros_like(_t(x))


def eye(n, dtype=None):
    return torch.eye(n, dtype=_dtype(dtype))


def arange(start, stop=None, step=1, dtype=None):
    if stop is None:
        start, stop = 0, start
    is_float = any(isinstance(a, float) for a in (start, stop, step))
    if dtype is None:
        dtype = torch.get_default_dtype() if is_float else torch.int64
    return torch.arange(start, stop, step, dtype=_dtype(dtype))


def linspace(start, stop, num=50, endpoint=True):
    assert endpoint
    return torch.linspace(start, stop, num)


def sqrt(x):
    return torch.sqrt(_t(x))


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def abs(x):  # noqa: A001
    return torch.abs(_t(x))


def square(x):
    return torch.square(_t(x))


def maximum(a, b):
    a, b = _t(a), _t(b)
    if a.ndim == 0 and b.ndim > 0:
        a = a.to(b.dtype)
    if b.ndim == 0 and a.ndim > 0:
        b = b.to(a.dtype)
    return torch.maximum(a, b)


def dot(a, b):
    a, b = _t(a), _t(b)
    if a.ndim == 1 and b.ndim == 1:
        return torch.sum(a * b)
    return a @ b


def outer(a, b):
    return torch.outer(_t(a), _t(b))


def tril(m, k=0):
    return torch.tril(_t(m), k)


def triu(m, k=0):
    return torch.triu(_t(m), k)


def diag(m, k=0):
    return torch.diag(_t(m), k)


def concatenate(xs, axis=0):
    return torch.cat([_t(x) for x in xs], dim=axis)


def stack(xs, axis=0):
    return torch.stack([_t(x) for x in xs], dim=axis)


def reshape(x, shape):
    return torch.reshape(_t(x), shape)


def flip(x, axis=None):
    x = _t(x)
    return torch.flip(x, dims=tuple(range(x.ndim)) if axis is None else (axis,))


def sum(x, axis=None):  # noqa: A001
    return torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis)


def mean(x, axis=None):
    return torch.mean(_t(x)) if axis is None else torch.mean(_t(x), dim=axis)


def std(x, axis=None):
    x = _t(x)
    return torch.std(x, unbiased=False) if axis is None else torch.std(x, dim=axis, unbiased=False)


def pad(x, width, mode="constant", constant_values=0.0):
    x = _t(x)
    assert x.ndim == 2 and width == 1
    if mode == "edge":
        return torch.nn.functional.pad(x[None, None], (1, 1, 1, 1), mode="replicate")[0, 0]
    return torch.nn.functional.pad(x, (1, 1, 1, 1), value=constant_values)


def allclose(a, b, rtol=1e-5, atol=1e-8):
    a, b = _t(a), _t(b)
    return bool(torch.allclose(a, b.to(a.dtype), rtol=rtol, atol=atol))


def all(x):  # noqa: A001
    return bool(torch.all(_t(x)))


def where(c, a, b):
    return torch.where(c, _t(a), _t(b))


def diff(x):
    return torch.diff(_t(x))


def meshgrid(*xs):
    return torch.meshgrid(*xs, indexing="xy")


class finfo:
    def __init__(self, x):
        dt = x.dtype if isinstance(x, torch.Tensor) else _dtype(x)
        self.eps = torch.finfo(dt).eps


def dtype(x):
    return x.dtype if isinstance(x, torch.Tensor) else _dtype(x)


def argmax(x, axis=None):
    return torch.argmax(_t(x)) if axis is None else torch.argmax(_t(x), dim=axis)


def argsort(x):
    return torch.argsort(_t(x), stable=True)


def logical_and(a, b):
    return torch.logical_and(torch.as_tensor(a), torch.as_tensor(b))


def logical_or(a, b):
    return torch.logical_or(torch.as_tensor(a), torch.as_tensor(b))


def amax(x, axis=None):
    return torch.amax(_t(x)) if axis is None else torch.amax(_t(x), dim=axis)


def minimum(a, b):
    return torch.minimum(_t(a), _t(b, _t(a).dtype) if not isinstance(b, torch.Tensor) else b)
