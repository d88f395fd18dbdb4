"""Torch-backed stand-in for `jax` (golden-vector generation only) — see `_core.py`."""

from jax import _core
from jax._core import (  # noqa: F401
    Array,
    checkpoint,
    closure_convert,
    config,
    custom_vjp,
    grad,
    jacfwd,
    jacrev,
    jit,
    value_and_grad,
    vjp,
    vmap,
)
from jax import numpy  # noqa: F401,E402
from jax import flatten_util, lax, nn, random, scipy, tree_util  # noqa: F401,E402
