import torch

from jax._core import _as_tensor, tree_flatten, tree_unflatten


def ravel_pytree(tree):
    leaves, treedef = tree_flatten(tree)
    leaves = [_as_tensor(x) for x in leaves]
    shapes = [x.shape for x in leaves]
    sizes = [x.numel() for x in leaves]
    flat = torch.cat([x.reshape(-1) for x in leaves]) if leaves else torch.zeros(0)

    def unravel(f):
        out, pos = [], 0
        for shp, sz in zip(shapes, sizes):
            out.append(f[pos : pos + sz].reshape(shp))
            pos += sz
        return tree_unflatten(treedef, out)

    return flat, unravel
