import torch


def expm(m, max_squarings=16):
    return torch.linalg.matrix_exp(m)
