"""`jax.lax` subset as eager Python loops — see `jax/_core.py`."""

import torch

from jax._core import _as_tensor, tree_flatten, tree_map, tree_unflatten


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(lower, upper):
        val = body_fun(i, val)
    return val


def scan(f, init, xs, reverse=False, length=None):
    leaves, _ = tree_flatten(xs)
    n = len(leaves[0]) if leaves else length
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        x_i = tree_map(lambda leaf: leaf[i], xs)
        carry, y = f(carry, x_i)
        ys[i] = y
    y_leaves0, y_def = tree_flatten(ys[0]) if n else ([], None)
    stacked = [
        torch.stack([_as_tensor(tree_flatten(y)[0][k]) for y in ys])
        for k in range(len(y_leaves0))
    ]
    return carry, (tree_unflatten(y_def, stacked) if n else ())


def map(f, xs):  # noqa: A001
    _, ys = scan(lambda c, x: (c, f(x)), None, xs)
    return ys


def stop_gradient(x):
    return tree_map(lambda t: _as_tensor(t).detach(), x)


def select(pred, on_true, on_false):
    return torch.where(pred, on_true, on_false)
