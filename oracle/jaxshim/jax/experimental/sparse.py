import torch


class BCOO:
    """COO matrix whose `@` sums duplicates (scatter-add), as `jax.experimental.sparse.BCOO`."""

    def __init__(self, args, shape):
        self.data, self.indices = args
        self.shape = tuple(shape)

    def __matmul__(self, x):
        row, col = self.indices[:, 0].long(), self.indices[:, 1].long()
        out = torch.zeros(self.shape[0], dtype=torch.result_type(self.data, x))
        return out.index_add(0, row, self.data * x[col])
