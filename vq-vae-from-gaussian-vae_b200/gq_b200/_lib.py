"""ctypes binding of libgq_b200.so -- the C ABI declared in include/gq_b200.h.

There is deliberately no fallback: if the shared object is missing or does not export the
expected symbols the import fails loudly (the product path is CUDA-only).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32,
                    c_uint64, c_void_p)
from pathlib import Path

# GQ_B200_LIB points the binding at another build of the same library (instrumented diagnostic
# builds, scripts/tc_prof.py); it is still a CUDA build of this ABI, never a fallback.
LIB_PATH = Path(os.environ.get("GQ_B200_LIB") or (Path(__file__).resolve().parent / "libgq_b200.so"))

GQ_OK = 0
GQ_LAYOUT_BLC = 0
GQ_LAYOUT_BCHW = 1
GQ_MODE_AUTO, GQ_MODE_FAST, GQ_MODE_ROBUST = 0, 1, 2
ABI_VERSION = 4

# name -> (restype, argtypes); mirrors include/gq_b200.h one to one
PROTOTYPES = {
    "gq_abi_version": (c_int, []),
    "gq_last_error": (c_char_p, []),
    "gq_launch_count": (c_uint64, []),
    "gq_scores_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                              c_double, c_void_p]),
    "gq_codebook_create": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p,
                                   POINTER(c_void_p)]),
    "gq_codebook_destroy": (None, [c_void_p]),
    "gq_codebook_set_mode": (c_int, [c_void_p, c_int]),
    "gq_codebook_set_usage_histogram": (c_int, [c_void_p, c_void_p]),
    "gq_codebook_size": (c_int64, [c_void_p]),
    "gq_codebook_group": (c_int, [c_void_p]),
    "gq_codebook_data": (c_void_p, [c_void_p]),
    "gq_quantize_workspace_bytes": (c_int64, [c_void_p, c_int64]),
    "gq_quantize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                            c_void_p, c_int64, c_void_p, c_void_p]),
    "gq_quantize_latent_workspace_bytes": (c_int64, [c_void_p, c_int64, c_int64, c_int]),
    "gq_quantize_latent": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_float, c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "gq_quantize_exact": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_void_p]),
    "gq_dequant": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p,
                           c_void_p]),
    "gq_index_width": (c_int, [c_int64]),
    "gq_pack_indices": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "gq_peer_gather_indices": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "gq_peer_gather_indices_sync": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64,
                                            c_uint32, c_void_p, c_void_p]),
    "gq_peer_kl_fold": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_uint32, c_double, c_double, c_double,
                                c_void_p, c_void_p, c_void_p]),
    "gq_unpack_indices": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "gq_dequant_packed": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p,
                                  c_void_p]),
    "gq_index_histogram": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "gq_usage_entropy": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gq_kl_workspace_bytes": (c_int64, [c_int64, c_int64, c_int]),
    "gq_kl_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                          c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_float, c_float,
                          c_void_p, c_int64, c_void_p]),
    "gq_kl_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                          c_int64, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p]),
    "gq_kl_fold": (c_int, [c_void_p, c_int, c_double, c_double, c_double, c_void_p, c_void_p]),
    # measurement / diagnostics (no reference counterpart)
    "gq_profile_enable": (None, [c_int]),
    "gq_profile_read": (c_int, [POINTER(c_double), POINTER(c_int)]),
    "gq_set_window_max": (None, [c_float]),
    "gq_set_chained_launch": (None, [c_int]),
    "gq_debug_tc_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_void_p]),
    "gq_debug_plan": (None, [c_int64, c_int64, c_int, c_int, POINTER(c_int64)]),
}

_lib = None


class GqError(RuntimeError):
    """A libgq_b200 entry point returned a non-zero status (the analogue of the reference's
    TORCH_CHECK failures, gq_cuda.cu:91-101, which surface as RuntimeError)."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python vq-vae-from-gaussian-vae_b200/build_ext.py` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        l = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if l.gq_abi_version() != ABI_VERSION:
            raise ImportError(f"libgq_b200 ABI {l.gq_abi_version()} != expected {ABI_VERSION}")
        if os.environ.get("GQ_B200_CHAINED_LAUNCH", "1") == "0":     # diagnostics: plain stream order
            l.gq_set_chained_launch(0)
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != GQ_OK:
        msg = lib().gq_last_error()
        raise GqError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
