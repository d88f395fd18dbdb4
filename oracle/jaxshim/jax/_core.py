"""A small torch-backed stand-in for the parts of the `jax` API that the reference's
hot-path modules use.  TEST INFRASTRUCTURE (golden-vector generation only).

Purpose: JAX is not installed in this image, so the reference
(`/root/reference/src/matfree_extensions/*.py`) cannot be imported as-is.  With this
package first on `sys.path`, `import jax` resolves here and the reference's
UNMODIFIED sources run eagerly on CPU torch tensors; `jax.vjp`/`jax.grad` map to
`torch.autograd`, `jax.custom_vjp` to `torch.autograd.Function`, `lax.fori_loop` /
`lax.scan` / `vmap` to Python loops.  Not a general JAX replacement: no tracing, no
XLA, no threefry PRNG (random numbers come from NumPy and are never part of a golden
contract — probes and inputs are stored in the fixtures explicitly).
"""

from __future__ import annotations

import functools
import warnings

import numpy as _np
import torch

warnings.filterwarnings("ignore", message="The use of `x.T` on tensors")

Array = torch.Tensor

# ---------------------------------------------------------------------------
# `x.at[idx].set(v)` on tensors (functional update, differentiable)
# ---------------------------------------------------------------------------


class _AtIndexer:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtSetter(self._arr, idx)


class _AtSetter:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def _in_bounds(self):
        # JAX drops out-of-bounds scatter updates (arnoldi.py:98 relies on it)
        idx = self._idx if isinstance(self._idx, tuple) else (self._idx,)
        for i, size in zip(idx, self._arr.shape):
            if isinstance(i, (int, _np.integer)) and not (-size <= i < size):
                return False
        return True

    def set(self, value):
        if not self._in_bounds():
            return self._arr
        out = self._arr.clone()
        out[self._idx] = torch.as_tensor(value, dtype=out.dtype)
        return out

    def add(self, value):
        if not self._in_bounds():
            return self._arr
        out = self._arr.clone()
        out[self._idx] = out[self._idx] + value
        return out


torch.Tensor.at = property(lambda self: _AtIndexer(self))
# JAX arrays are immutable: `v /= length` (arnoldi.py:80) rebinds, it never mutates
torch.Tensor.__itruediv__ = lambda self, other: self / other
torch.Tensor.__imul__ = lambda self, other: self * other
torch.Tensor.__iadd__ = lambda self, other: self + other
torch.Tensor.__isub__ = lambda self, other: self - other
torch.Tensor.block_until_ready = lambda self: self

# ---------------------------------------------------------------------------
# pytrees
# ---------------------------------------------------------------------------


class _Leaf:
    pass


_LEAF = _Leaf()


def tree_flatten(tree):
    leaves = []

    def rec(t):
        if t is None:
            return ("none",)
        if isinstance(t, Partial):
            return ("partial", t.func, rec(tuple(t.args)), rec(dict(t.keywords)))
        if isinstance(t, tuple):
            return ("tuple", [rec(x) for x in t])
        if isinstance(t, list):
            return ("list", [rec(x) for x in t])
        if isinstance(t, dict):
            keys = sorted(t.keys())
            return ("dict", keys, [rec(t[k]) for k in keys])
        leaves.append(t)
        return _LEAF

    treedef = rec(tree)
    return leaves, treedef


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        if d is _LEAF:
            return next(it)
        kind = d[0]
        if kind == "none":
            return None
        if kind == "tuple":
            return tuple(rec(x) for x in d[1])
        if kind == "list":
            return [rec(x) for x in d[1]]
        if kind == "dict":
            return {k: rec(x) for k, x in zip(d[1], d[2])}
        if kind == "partial":
            return Partial(d[1], *rec(d[2]), **rec(d[3]))
        raise TypeError(kind)

    return rec(treedef)


def tree_map(f, tree, *rest):
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


class Partial(functools.partial):
    """`jax.tree_util.Partial`: a partial that is also a pytree."""


def _as_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(x)


# ---------------------------------------------------------------------------
# autodiff
# ---------------------------------------------------------------------------


def vjp(fun, *primals, has_aux=False):
    flat, treedef = tree_flatten(primals)
    flat = [_as_tensor(x).detach().clone().requires_grad_(True) for x in flat]
    with torch.enable_grad():
        out = fun(*tree_unflatten(treedef, flat))
        if has_aux:
            out, aux = out
        out_flat, out_def = tree_flatten(out)
        out_flat = [_as_tensor(o) for o in out_flat]

    def pullback(cot):
        cot_flat, _ = tree_flatten(cot)
        pairs = [
            (o, torch.as_tensor(c, dtype=o.dtype).reshape(o.shape))
            for o, c in zip(out_flat, cot_flat)
            if o.requires_grad
        ]
        if pairs:
            grads = torch.autograd.grad(
                [o for o, _ in pairs], flat, [c for _, c in pairs],
                allow_unused=True, retain_graph=True,
            )  # fmt: skip
        else:
            grads = [None] * len(flat)
        grads = [torch.zeros_like(x) if g is None else g for g, x in zip(grads, flat)]
        return tree_unflatten(treedef, grads)

    detached = tree_unflatten(out_def, [o.detach() for o in out_flat])
    if has_aux:
        return detached, pullback, aux
    return detached, pullback


def value_and_grad(fun, argnums=0, has_aux=False):
    def wrapped(*args):
        single = isinstance(argnums, int)
        nums = (argnums,) if single else tuple(argnums)

        def partial_fun(*diff):
            full = list(args)
            for i, d in zip(nums, diff):
                full[i] = d
            return fun(*full)

        res = vjp(partial_fun, *[args[i] for i in nums], has_aux=has_aux)
        out, pull = res[0], res[1]
        grads = pull(torch.ones_like(out))
        grads = grads[0] if single else grads
        if has_aux:
            return (out, res[2]), grads
        return out, grads

    return wrapped


def grad(fun, argnums=0, has_aux=False):
    vg = value_and_grad(fun, argnums=argnums, has_aux=has_aux)

    def wrapped(*args):
        return vg(*args)[1]

    return wrapped


def jacfwd(fun):
    """Only used on scalar->scalar functions (`lanczos.py:112`)."""

    def wrapped(x):
        _, pull = vjp(fun, x)
        return pull(torch.ones_like(_as_tensor(x)))[0]

    return wrapped


jacrev = jacfwd


class custom_vjp:
    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        self.nondiff_argnums = tuple(nondiff_argnums)
        self.fwd = self.bwd = None
        self._depth = 0
        functools.update_wrapper(self, fun)

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *args):
        if self._depth > 0 or self.fwd is None:
            # a call from inside the function's own fwd rule (arnoldi.py:30) is not being
            # differentiated at that level: JAX evaluates the primal function
            return self.fun(*args)
        nondiff = [args[i] for i in self.nondiff_argnums]
        diff = tuple(a for i, a in enumerate(args) if i not in self.nondiff_argnums)
        flat_in, def_in = tree_flatten(diff)
        flat_in = [_as_tensor(x) for x in flat_in]
        fwd, bwd = self.fwd, self.bwd
        box = {}
        outer = self

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *flat):
                outer._depth += 1
                try:
                    with torch.no_grad():
                        out, res = fwd(*nondiff, *tree_unflatten(def_in, list(flat)))
                finally:
                    outer._depth -= 1
                out_flat, out_def = tree_flatten(out)
                ctx.res = res
                box["out_def"] = out_def
                return tuple(_as_tensor(o) for o in out_flat)

            @staticmethod
            def backward(ctx, *cot_flat):
                cot = tree_unflatten(box["out_def"], list(cot_flat))
                with torch.enable_grad():
                    grads = bwd(*nondiff, ctx.res, cot)
                g_flat, _ = tree_flatten(tuple(grads))
                return tuple(_as_tensor(g).detach() for g in g_flat)

        out_flat = _Fn.apply(*flat_in)
        return tree_unflatten(box["out_def"], list(out_flat))


def closure_convert(fun, *example_args):
    return fun, []


def jit(fun=None, **_kwargs):
    if fun is None:
        return lambda f: f
    return fun


def checkpoint(fun, **_kwargs):
    return fun


def _take(x, axis, i):
    return x if axis is None else torch.select(_as_tensor(x), axis, i)


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        size = None
        for a, ax in zip(args, axes):
            if ax is not None:
                leaves, _ = tree_flatten(a)
                size = _as_tensor(leaves[0]).shape[ax]
                break
        outs = []
        for i in range(size):
            sl = [
                a if ax is None else tree_map(lambda t, ax=ax: _take(t, ax, i), a)
                for a, ax in zip(args, axes)
            ]
            outs.append(fun(*sl))
        flat0, out_def = tree_flatten(outs[0])
        stacked = []
        for k in range(len(flat0)):
            items = [_as_tensor(tree_flatten(o)[0][k]) for o in outs]
            ax = out_axes
            if ax < 0:
                ax = items[0].ndim + 1 + ax
            stacked.append(torch.stack(items, dim=ax))
        return tree_unflatten(out_def, stacked)

    return mapped


class _Config:
    x64 = False

    def update(self, name, value):
        if name == "jax_enable_x64":
            self.x64 = bool(value)
            torch.set_default_dtype(torch.float64 if value else torch.float32)


config = _Config()
