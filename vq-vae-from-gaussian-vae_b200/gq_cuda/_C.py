"""Operator registration -- the counterpart of the reference's `_C` module whose import runs the
TORCH_LIBRARY static initialisers (gq_cuda/csrc/gq_cuda.cpp:8-31, csrc/cuda/gq_cuda.cu:120-122).

Schema and dispatch key are identical: `extension_cpp::gq(Tensor a, Tensor b, Tensor c,
Tensor(a!) out, int d, int e, int f, float g) -> ()`, CUDA only (CPU tensors raise
NotImplementedError from the dispatcher, as with the reference)."""
import torch

from gq_b200 import ops as _ops

_SCHEMA = "gq(Tensor a, Tensor b, Tensor c, Tensor(a!) out, int d, int e, int f, float g) -> ()"

_lib = torch.library.Library("extension_cpp", "DEF")
_lib.define(_SCHEMA)


def _gq_cuda(a, b, c, out, d, e, f, g):
    _ops.scores_legacy_(a, b, c, out, d, e, f, g)


_lib.impl("gq", _gq_cuda, "CUDA")
