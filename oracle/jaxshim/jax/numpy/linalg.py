import torch


def norm(x):
    return torch.linalg.norm(x)


def eigh(m):
    # jnp.linalg.eigh symmetrises its input by default
    return torch.linalg.eigh(0.5 * (m + m.T.conj()))


def solve(a, b):
    return torch.linalg.solve(a, b)
