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

_INSTALLED = {}


def install(module: str = "pit.quantization.gaussian"):
    """Bind this package's classes over the reference's at `module` -- no edit to the reference tree.

    `instantiate_from_config` (pit/util.py:45-62) resolves `target: pit.quantization.gaussian.GaussianQuantRegularizer`
    with importlib + getattr at call time, so rebinding the module attributes once (after the reference tree is on
    sys.path, before the model is built) makes every existing config construct the B200 regularizer with its own
    parameters unchanged.  Rebinds GaussianQuantRegularizer, GaussianQuantRegularizer2 and prior_samples, and sets
    HAS_GQ_CUDA (gaussian.py:9-13).  Returns the names it rebound; `uninstall()` restores the originals."""
    import importlib
    mod = importlib.import_module(module)
    names = {"GaussianQuantRegularizer": GaussianQuantRegularizer, "GaussianQuantRegularizer2": GaussianQuantRegularizer2,
             "prior_samples": prior_samples, "HAS_GQ_CUDA": True}
    saved = _INSTALLED.setdefault(module, {})
    for k, v in names.items():
        if k not in saved:
            saved[k] = getattr(mod, k, None)
        setattr(mod, k, v)
    return sorted(names)


def uninstall(module: str = "pit.quantization.gaussian") -> None:
    import importlib
    saved = _INSTALLED.pop(module, None)
    if not saved:
        return
    mod = importlib.import_module(module)
    for k, v in saved.items():
        if v is None:
            if hasattr(mod, k):
                delattr(mod, k)
        else:
            setattr(mod, k, v)


__all__ = [
    "install", "uninstall", "ops", "Codebook", "quantize_rows", "quantize_latent", "dequant", "kl_penalty", "scores_legacy_", "launch_count",
    "pack_indices", "unpack_indices", "packed_dtype", "index_histogram", "usage_entropy",
    "GaussianQuantRegularizer", "GaussianQuantRegularizer2", "prior_samples",
]
