import torch


def expm(m, max_squarings=16):
    return torch.linalg.matrix_exp(m)


def cho_factor(a, lower=False):
    return torch.linalg.cholesky(a), True


def cho_solve(c_and_lower, b):
    c, _lower = c_and_lower
    return torch.cholesky_solve(b.reshape(len(b), -1), c).reshape(b.shape)


def solve_triangular(a, b, lower=False, trans=False):
    assert not trans
    return torch.linalg.solve_triangular(a, b.reshape(len(b), -1), upper=not lower).reshape(b.shape)
