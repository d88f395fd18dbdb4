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


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def custom_linear_solve(matvec, b, solve, transpose_solve=None, symmetric=False, has_aux=False):
    """`jax.lax.custom_linear_solve`: the primal value comes from `solve(matvec, b)` (never
    differentiated, exactly as in JAX), the derivative from the implicit function theorem:
    `dx = A^{-1} (db - dA x)`, with `A^{-1}` applied by `solve` again (symmetric systems only)."""
    assert symmetric and transpose_solve is None

    def run(rhs):
        with torch.no_grad():
            out = solve(matvec, rhs.detach())
        sol, aux = out if has_aux else (out, None)
        return sol.detach(), aux

    x, aux = run(_as_tensor(b))

    class _Inverse(torch.autograd.Function):  # r -> A^{-1} r as a LINEAR map with a solve-based transpose
        @staticmethod
        def forward(ctx, r):
            return run(r)[0]

        @staticmethod
        def backward(ctx, g):
            return run(g)[0]

    with torch.enable_grad():
        residual = _as_tensor(b) - matvec(x)  # carries the dependence on b and on matvec's closure
    # value: x + A^{-1} r - A^{-1} r == x exactly; derivative: A^{-1} (db - dA x)
    out = x + _Inverse.apply(residual) - _Inverse.apply(residual.detach()).detach()
    aux = tree_map(lambda t: _as_tensor(t).detach(), aux) if has_aux else None
    return (out, aux) if has_aux else out
