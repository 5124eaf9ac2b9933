"""CPU oracle for the GaussianQuant hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this package; the product (`gq_b200`, `gq_cuda`) never does.

ctypes front-end over oracle/gq_oracle.c (a C restatement of
pit/quantization/gaussian.py:76-98,121-178 and gq_cuda_extension/.../gq_cuda.cu:12-40).
Parity status: pinned against the unmodified reference via tests/golden/*.npz.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libgq_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile gq_oracle.c with gcc (seconds)."""
    src = _HERE / "gq_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.gq_oracle_version.restype = ctypes.c_int
        _lib.gq_oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().gq_oracle_num_threads())


def set_num_threads(n: int = 0) -> int:
    """Use n OpenMP threads (0 = every online core); returns the count now in effect."""
    lib().gq_oracle_set_num_threads(ctypes.c_int(n))
    return num_threads()


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a: np.ndarray, ty=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def normal_log_prob(codebook) -> np.ndarray:
    """gaussian.py:51-52."""
    cb = _f32(codebook)
    out = np.empty_like(cb)
    lib().gq_oracle_normal_log_prob(_p(cb), ctypes.c_int64(cb.size), _p(out))
    return out


def scores_torch(mu, std, log_std, codebook, nlp, beta: float = 1.0) -> np.ndarray:
    """[M,N] fp32 log-ratio matrix in the reference torch op order (gaussian.py:141-147)."""
    mu, std, log_std, cb, nlp = map(_f32, (mu, std, log_std, codebook, nlp))
    M, g = mu.shape
    N = cb.shape[0]
    out = np.empty((M, N), np.float32)
    rc = lib().gq_oracle_scores_torch(_p(mu), _p(std), _p(log_std), _p(cb), _p(nlp),
                                      ctypes.c_int64(M), ctypes.c_int64(N), ctypes.c_int(g),
                                      ctypes.c_float(beta), _p(out))
    if rc:
        raise RuntimeError(f"gq_oracle_scores_torch rc={rc}")
    return out


def quantize(mu, std, log_std, codebook, nlp, beta: float = 1.0):
    """Eval-branch quantizer rows -> (idx int64 [M], zhat [M,g], best [M], gap [M]).

    gaussian.py:134-150; `gap` is the fp32 top-1 minus top-2 score of the row."""
    mu, std, log_std, cb, nlp = map(_f32, (mu, std, log_std, codebook, nlp))
    M, g = mu.shape
    N = cb.shape[0]
    idx = np.empty(M, np.int64)
    zhat = np.empty((M, g), np.float32)
    best = np.empty(M, np.float32)
    gap = np.empty(M, np.float32)
    rc = lib().gq_oracle_quantize(_p(mu), _p(std), _p(log_std), _p(cb), _p(nlp),
                                  ctypes.c_int64(M), ctypes.c_int64(N), ctypes.c_int(g),
                                  ctypes.c_float(beta), _p(idx, ctypes.c_int64), _p(zhat),
                                  _p(best), _p(gap))
    if rc:
        raise RuntimeError(f"gq_oracle_quantize rc={rc}")
    return idx, zhat, best, gap


def scores_legacy(mu, std, noise, beta: float = 1.0, contract: bool = True) -> np.ndarray:
    """Legacy `gq_cuda` operator output (gq_cuda.cu:31-38)."""
    mu, std, noise = map(_f32, (mu, std, noise))
    M, dim = mu.shape
    N = noise.shape[0]
    out = np.empty((M, N), np.float32)
    lib().gq_oracle_scores_legacy(_p(mu), _p(std), _p(noise), _p(out), ctypes.c_int64(dim),
                                  ctypes.c_int64(M), ctypes.c_int64(N), ctypes.c_double(beta),
                                  ctypes.c_int(1 if contract else 0))
    return out


def gather(idx, codebook) -> np.ndarray:
    """dequant rows (gaussian.py:171-173)."""
    idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int64)).reshape(-1)
    cb = _f32(codebook)
    out = np.empty((idx.size, cb.shape[1]), np.float32)
    rc = lib().gq_oracle_gather(_p(idx, ctypes.c_int64), _p(cb), ctypes.c_int64(idx.size),
                                ctypes.c_int64(cb.shape[0]), ctypes.c_int(cb.shape[1]), _p(out))
    if rc:
        raise IndexError(f"gq_oracle_gather rc={rc}")
    return out


def kl(z_blc, group: int, n_samples_log2: int, tolerance: float = 0.5, lam: float = 1.0,
       lam_min: float = 1.0, lam_max: float = 1.0, logvar_range=(-30.0, 20.0)):
    """Train-branch KL penalty on a blc latent (gaussian.py:76-98).

    Returns dict(kl_loss, bits_mean, bits_min, bits_max, kl2 [B,L,ncb])."""
    z = _f32(z_blc)
    B, L, C2 = z.shape
    C = C2 // 2
    ncb = C // group
    stats = np.zeros(5, np.float64)
    kl2 = np.empty((B, L, ncb), np.float32)
    rc = lib().gq_oracle_kl(_p(z), ctypes.c_int64(B), ctypes.c_int64(L), ctypes.c_int(C),
                            ctypes.c_int(group), ctypes.c_float(logvar_range[0]),
                            ctypes.c_float(logvar_range[1]), ctypes.c_float(n_samples_log2),
                            ctypes.c_float(tolerance), ctypes.c_float(lam_min),
                            ctypes.c_float(lam_max), _p(stats, ctypes.c_double), _p(kl2))
    if rc:
        raise RuntimeError(f"gq_oracle_kl rc={rc}")
    return {
        "kl_loss": lam * stats[3] / B,
        "bits_mean": stats[0] / stats[4],
        "bits_min": stats[1],
        "bits_max": stats[2],
        "kl2": kl2,
    }


# ---- index stream after the quantizer (numpy restatements; test infrastructure like the rest) ----------
def index_histogram(idx, n_samples, hist=None):
    """eval.py:127,152-154: all_hist (zeros(N)) += torch.histogram(indices.float(), torch.range(0, N))[0].
    Bin i counts the indices equal to i; values outside [0, N) are dropped.  int64 counts."""
    import numpy as _np
    idx = _np.asarray(idx).reshape(-1)
    ok = (idx >= 0) & (idx < n_samples)
    h = _np.bincount(idx[ok].astype(_np.int64), minlength=int(n_samples)).astype(_np.int64)
    if hist is None:
        return h
    hist += h
    return hist


def usage_entropy(hist):
    """cal_ent, eval.py:137-141, in the reference's fp32 tensor arithmetic:
    usage = 1 - sum(hist == 0)/N; p = hist / sum(hist); ent = -sum(p * log2(p + 1e-5))."""
    import numpy as _np
    h = _np.asarray(hist).astype(_np.float32)
    unused = _np.float32((h == 0).astype(_np.float32).sum(dtype=_np.float32)) / _np.float32(h.shape[0])
    p = h / h.sum(dtype=_np.float32)
    ent = -(p * _np.log2(p + _np.float32(1e-5))).sum(dtype=_np.float32)
    return float(_np.float32(1) - unused), float(ent)


def packed_dtype(n_samples):
    import numpy as _np
    return _np.uint16 if n_samples <= 65536 else _np.uint32


def pack_indices(idx, n_samples):
    import numpy as _np
    return _np.asarray(idx).astype(packed_dtype(n_samples))


def unpack_indices(packed):
    import numpy as _np
    return _np.asarray(packed).astype(_np.int64)
