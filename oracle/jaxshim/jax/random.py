"""NumPy-seeded random numbers.  NOT JAX's threefry stream: values drawn here are
never part of a golden contract (fixtures store every input explicitly)."""

import numpy as _np
import torch


def PRNGKey(seed):
    return torch.tensor([0, int(seed)], dtype=torch.int64)


def _rng(key):
    return _np.random.default_rng([int(k) for k in key.tolist()])


def split(key, num=2):
    rng = _rng(key)
    return torch.as_tensor(rng.integers(0, 2**31 - 1, size=(num, 2)), dtype=torch.int64)


def normal(key, shape=(), dtype=None):
    out = torch.as_tensor(_rng(key).standard_normal(shape))
    return out.to(dtype or torch.get_default_dtype())


def uniform(key, shape=(), dtype=None):
    out = torch.as_tensor(_rng(key).uniform(size=shape))
    return out.to(dtype or torch.get_default_dtype())


def rademacher(key, shape=(), dtype=None):
    out = torch.as_tensor(_rng(key).integers(0, 2, size=shape) * 2.0 - 1.0)
    return out.to(dtype or torch.get_default_dtype())
