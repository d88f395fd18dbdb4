"""B200-native Lanczos/Arnoldi adjoints: the hot path of pnkraemer/experiments-lanczos-adjoints
(`matfree_extensions.{arnoldi,lanczos,hutchinson}` + its matvec back-ends) behind the
reference's own function API.  All device work goes through libb200lanczos.so (hand-written
sm_100a CUDA, C ABI in `include/b200_lanczos.h`); there is no CPU fallback.
"""

from experiments_lanczos_adjoints_b200 import arnoldi, cg, gp, hutchinson, lanczos, low_rank, operators, pde  # noqa: F401
from experiments_lanczos_adjoints_b200._lib import BLError  # noqa: F401
from experiments_lanczos_adjoints_b200.device import (  # noqa: F401
    DeviceArray,
    Event,
    Stream,
    asarray,
    default_stream,
    device_count,
    empty,
    empty_cache,
    launch_count,
    set_blocks_per_sm,
    set_device,
    sm_count,
    synchronize,
    zeros,
)


def vjp(fun, *primals):
    """`jax.vjp(fun, *primals)` for the function objects of this package:
    returns `(outputs, pullback)`; `pullback(cotangents) -> (d primal_0, d primal_1, ...)`."""
    return fun.vjp(*primals)


def value_and_grad(fun, argnums=0):
    """`jax.value_and_grad(fun, argnums)` for integrands / estimators of this package."""

    def wrapped(*args):
        value, grads = fun.value_and_grad(*args)
        full = (None, *grads) if len(grads) == len(args) - 1 else tuple(grads)
        if isinstance(argnums, int):
            return value, full[argnums]
        return value, tuple(full[i] for i in argnums)

    return wrapped
