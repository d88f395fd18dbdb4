import torch


def relu(x):
    return torch.relu(x)


def softplus(x):
    return torch.nn.functional.softplus(x)
