"""gq_b200 -- B200-native (sm_100a) GaussianQuant hot path behind the reference's API.

    from gq_b200 import GaussianQuantRegularizer        # drop-in module
    import gq_cuda; gq_cuda.ops.gq_cuda(...)            # drop-in legacy operator

Importing this package loads libgq_b200.so; there is no fallback if it is missing.
"""
from . import _lib

_lib.lib()  # fail loudly at import time if the CUDA library is absent or stale

from . import ops  # noqa: E402
from .ops import (Codebook, dequant, index_histogram, kl_penalty, launch_count, pack_indices,  # noqa: E402
                  packed_dtype, quantize_latent, quantize_rows, scores_legacy_, unpack_indices, usage_entropy)
from .regularizer import GaussianQuantRegularizer, GaussianQuantRegularizer2, prior_samples  # noqa: E402

__all__ = [
    "ops", "Codebook", "quantize_rows", "quantize_latent", "dequant", "kl_penalty", "scores_legacy_", "launch_count",
    "pack_indices", "unpack_indices", "packed_dtype", "index_histogram", "usage_entropy",
    "GaussianQuantRegularizer", "GaussianQuantRegularizer2", "prior_samples",
]
