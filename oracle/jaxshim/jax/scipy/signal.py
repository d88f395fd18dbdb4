import torch


def convolve2d(in1, in2, mode="full"):
    """True 2-D convolution, `mode="valid"`, small kernel `in1` over `in2`
    (the call in `/root/reference/src/matfree_extensions/util/pde_util.py:137`)."""
    assert mode == "valid"
    kernel = torch.flip(in1, dims=(0, 1)).to(in2.dtype)
    return torch.nn.functional.conv2d(in2[None, None], kernel[None, None])[0, 0]
