import torch
from torch import Tensor

__all__ = ["gq_cuda"]


def gq_cuda(a: Tensor, b: Tensor, c: Tensor, out: Tensor, d: int, e: int, f: int, g: float) -> None:
    """Same signature as gq_cuda_extension/gq_cuda/ops.py:7-8: mu[M,dim], std[M,dim],
    noise[N,dim], out[M,N] (filled in place), dim, M, N, beta."""
    torch.ops.extension_cpp.gq.default(a, b, c, out, d, e, f, g)
