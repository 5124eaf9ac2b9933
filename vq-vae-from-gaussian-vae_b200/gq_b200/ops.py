"""Tensor-level operators over the libgq_b200 C ABI.

PyTorch is plumbing here: it owns device memory and the current stream; every operator
below validates its tensors the way the reference wrapper does
(gq_cuda_extension/gq_cuda/csrc/cuda/gq_cuda.cu:91-101), then hands raw device pointers to
the C entry point on `torch.cuda.current_stream()`.  Nothing synchronises unless a caller opts into
`check_range=True` (dequant / pack_indices / index_histogram then read an error flag back, one blocking D2H).
A `Codebook` caches one scratch workspace and its device-side counters: use a handle (and the module that
owns it) from one stream at a time.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import GQ_LAYOUT_BCHW, GQ_LAYOUT_BLC, check

__all__ = [
    "Codebook", "scores_legacy_", "quantize_rows", "quantize_latent", "dequant", "kl_penalty", "launch_count",
    "GQ_LAYOUT_BLC", "GQ_LAYOUT_BCHW",
]

_SUPPORTED_GROUPS = (1, 2, 4, 8, 16, 32)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require_cuda_f32(name: str, t: torch.Tensor) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        # the reference registers no CPU kernel: calling with CPU tensors raises
        raise NotImplementedError(f"{name} must be a CUDA tensor (no CPU implementation exists)")


def launch_count() -> int:
    """Kernels launched by libgq_b200 in this process so far."""
    return int(_lib.lib().gq_launch_count())


def scores_legacy_(mu: torch.Tensor, std: torch.Tensor, codebook: torch.Tensor, out: torch.Tensor,
                   dim: int, M: int, N: int, beta: float) -> None:
    """Legacy operator `extension_cpp::gq` (gq_cuda.cu:77-118): fills out[M,N] in place."""
    for n, t in (("a", mu), ("b", std), ("c", codebook), ("out", out)):
        _require_cuda_f32(n, t)
    if mu.shape != std.shape:
        raise RuntimeError("mu_q.sizes() == std_q.sizes() check failed")
    if not out.is_contiguous():
        raise RuntimeError("result.is_contiguous() check failed")
    if mu.numel() != M * dim or codebook.numel() != N * dim or out.numel() != M * N:
        raise RuntimeError(f"shape mismatch: mu {tuple(mu.shape)}, codebook {tuple(codebook.shape)}, "
                           f"out {tuple(out.shape)} vs dim={dim} b={M} n_samples={N}")
    mu_c, std_c, cb_c = mu.contiguous(), std.contiguous(), codebook.contiguous()
    with torch.cuda.device(out.device):
        check(_lib.lib().gq_scores_f32(_ptr(mu_c), _ptr(std_c), _ptr(cb_c), _ptr(out), dim, M, N,
                                       float(beta), _stream()), "gq_scores_f32")


class Codebook:
    """Device-resident prepared codebook (gq_codebook_t).

    `samples` is the fp32 [N, g] tensor produced by the host-side prior_samples()
    (gaussian.py:15-19, 50); `nlp` the module's normal_log_prob buffer (gaussian.py:51-52) or
    None to have it recomputed in the same op order."""

    def __init__(self, samples: torch.Tensor, nlp: Optional[torch.Tensor] = None, beta: float = 1.0):
        _require_cuda_f32("samples", samples)
        if samples.dim() != 2:
            raise RuntimeError("samples must be [N, group]")
        N, g = samples.shape
        if g not in _SUPPORTED_GROUPS:
            raise RuntimeError(f"group {g} unsupported (supported: {_SUPPORTED_GROUPS})")
        if nlp is not None:
            _require_cuda_f32("nlp", nlp)
            if nlp.shape != samples.shape:
                raise RuntimeError("nlp must have the shape of samples")
            nlp = nlp.contiguous()
        samples = samples.contiguous()
        self.N, self.g, self.beta, self.device = int(N), int(g), float(beta), samples.device
        self._handle = _lib.ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.lib().gq_codebook_create(_ptr(samples), _ptr(nlp), N, g, float(beta), _stream(),
                                                _lib.ctypes.byref(self._handle)), "gq_codebook_create")
        self._ws = None  # cached quantize workspace (per handle, grown on demand)

    @property
    def handle(self):
        if not self._handle:
            raise RuntimeError("codebook handle was destroyed")
        return self._handle

    def set_mode(self, mode: str) -> None:
        """Search precision for groups 8 and 16 (gq_mode in gq_b200.h): "auto" (default: fast until many
        rows of a call had a too-wide fp16 window, then robust), "fast" (fp16 image, wide rows take the
        exact scan) or "robust" (split-precision image, ~1.4x the search time)."""
        modes = {"auto": _lib.GQ_MODE_AUTO, "fast": _lib.GQ_MODE_FAST, "robust": _lib.GQ_MODE_ROBUST}
        if mode not in modes:
            raise ValueError(f"mode must be one of {sorted(modes)}")
        check(_lib.lib().gq_codebook_set_mode(self.handle, modes[mode]), "gq_codebook_set_mode")

    def set_usage_histogram(self, hist: Optional[torch.Tensor]) -> None:
        """While set (int64 [N] CUDA tensor, accumulated in place; None to stop), every quantize_latent on
        this codebook also counts the indices it emits -- eval.py's all_hist, fused into the last pass."""
        if hist is not None:
            if hist.dtype != torch.int64 or hist.numel() != self.N or not hist.is_contiguous() or hist.device != self.device:
                raise RuntimeError("hist must be a contiguous int64 tensor of N bins on the codebook's device")
        self._usage_hist = hist          # keep the tensor alive while the library holds its address
        check(_lib.lib().gq_codebook_set_usage_histogram(self.handle, _ptr(hist)), "gq_codebook_set_usage_histogram")

    def workspace(self, M: int) -> torch.Tensor:
        need = int(_lib.lib().gq_quantize_workspace_bytes(self.handle, M))
        return self._grow(need)

    def latent_workspace(self, b: int, l: int, C: int) -> torch.Tensor:
        need = int(_lib.lib().gq_quantize_latent_workspace_bytes(self.handle, b, l, C))
        if need < 0:
            raise RuntimeError(f"latent with {C} channels does not fit group {self.g}")
        return self._grow(need)

    def _grow(self, need: int) -> torch.Tensor:
        # one cached scratch buffer per handle; calls on one stream are ordered, so reuse is safe
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def close(self) -> None:
        if getattr(self, "_handle", None):
            _lib.lib().gq_codebook_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def quantize_rows(cb: Codebook, mu: torch.Tensor, std: torch.Tensor,
                  log_std: Optional[torch.Tensor] = None, exact: bool = False,
                  stats: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-branch quantizer on rows (gaussian.py:124-150): mu, std [M, g] -> (idx int64 [M],
    zhat fp32 [M, g]).  `exact=True` runs the CUDA-core exact-order scan instead of the
    tensor-core search (same results, used as anchor)."""
    _require_cuda_f32("mu", mu)
    _require_cuda_f32("std", std)
    if mu.dim() != 2 or mu.shape != std.shape or mu.shape[1] != cb.g:
        raise RuntimeError(f"mu/std must both be [M, {cb.g}], got {tuple(mu.shape)} / {tuple(std.shape)}")
    if mu.device != cb.device:
        raise RuntimeError("rows and codebook live on different devices")
    mu, std = mu.contiguous(), std.contiguous()
    if log_std is not None:
        _require_cuda_f32("log_std", log_std)
        if log_std.shape != mu.shape:
            raise RuntimeError("log_std must have the shape of mu")
        log_std = log_std.contiguous()
    M = mu.shape[0]
    idx = torch.empty(M, dtype=torch.int64, device=mu.device)
    zhat = torch.empty(M, cb.g, dtype=torch.float32, device=mu.device)
    if M == 0:
        return idx, zhat
    with torch.cuda.device(mu.device):
        if exact:
            check(_lib.lib().gq_quantize_exact(cb.handle, _ptr(mu), _ptr(std), _ptr(log_std), M, _ptr(idx),
                                               _ptr(zhat), _stream()), "gq_quantize_exact")
        else:
            ws = cb.workspace(M)
            check(_lib.lib().gq_quantize(cb.handle, _ptr(mu), _ptr(std), _ptr(log_std), M, _ptr(idx),
                                         _ptr(zhat), _ptr(ws), ws.numel(), _ptr(stats), _stream()),
                  "gq_quantize")
    return idx, zhat


def quantize_latent(cb: Codebook, z: torch.Tensor, layout: int, eps: Optional[torch.Tensor] = None,
                    logvar_range=(-30.0, 20.0), stats: Optional[torch.Tensor] = None):
    """The whole eval branch from the raw latent (gaussian.py:62-80,121-159) in one C call.

    z fp32 [b,l,2C] (BLC) or [b,2C,h,w] (BCHW); eps fp32 noise shaped like zhat, or None.
    Returns (zhat, indices, zhat_noquant), all in z's layout (zhat_noquant None without eps)."""
    _require_cuda_f32("z", z)
    if z.device != cb.device:
        raise RuntimeError("latent and codebook live on different devices")
    z = z.contiguous()
    if layout == GQ_LAYOUT_BCHW:
        if z.dim() != 4:
            raise RuntimeError("bchw latent must be [b, 2C, h, w]")
        b, c2, h, w = z.shape
        l = h * w
    else:
        if z.dim() != 3:
            raise RuntimeError("blc latent must be [b, l, 2C]")
        b, l, c2 = z.shape
    C = c2 // 2
    if c2 != 2 * C or C % cb.g:
        raise RuntimeError(f"latent channels {c2} do not split into (mu | logvar) groups of {cb.g}")
    ncb = C // cb.g
    if layout == GQ_LAYOUT_BCHW:
        zhat = torch.empty(b, C, h, w, dtype=torch.float32, device=z.device)
        idx = torch.empty(b, ncb, h, w, dtype=torch.int64, device=z.device)
    else:
        zhat = torch.empty(b, l, C, dtype=torch.float32, device=z.device)
        idx = torch.empty(b, l, ncb, dtype=torch.int64, device=z.device)
    znq = None
    if eps is not None:
        _require_cuda_f32("eps", eps)
        if eps.shape != zhat.shape:
            raise RuntimeError(f"eps must be shaped like zhat {tuple(zhat.shape)}, got {tuple(eps.shape)}")
        eps = eps.contiguous()
        znq = torch.empty_like(zhat)
    if b * l == 0:
        return zhat, idx, znq
    ws = cb.latent_workspace(b, l, C)
    with torch.cuda.device(z.device):
        check(_lib.lib().gq_quantize_latent(cb.handle, _ptr(z), _ptr(eps), b, l, C, layout, float(logvar_range[0]),
                                            float(logvar_range[1]), _ptr(idx), _ptr(zhat), _ptr(znq), _ptr(ws),
                                            ws.numel(), _ptr(stats), _stream()), "gq_quantize_latent")
    return zhat, idx, znq


def dequant(cb: Codebook, indices: torch.Tensor, layout: int, check_range: bool = False) -> torch.Tensor:
    """gaussian.py:162-178.  indices int64 [b,l,ncb] (BLC) or [b,ncb,h,w] (BCHW)."""
    packed = indices.dtype in (torch.uint16, torch.int32)
    if indices.dtype != torch.int64 and not packed:
        raise RuntimeError(f"indices must be int64 (or a packed uint16/int32 stream), got {indices.dtype}")
    if packed and indices.dtype != packed_dtype(cb.N):
        raise RuntimeError(f"a {cb.N}-entry codebook packs to {packed_dtype(cb.N)}, got {indices.dtype}")
    if not indices.is_cuda:
        raise NotImplementedError("indices must be a CUDA tensor (no CPU implementation exists)")
    indices = indices.contiguous()
    if layout == GQ_LAYOUT_BCHW:
        if indices.dim() != 4:
            raise RuntimeError("bchw indices must be [b, ncb, h, w]")
        b, ncb, h, w = indices.shape
        l = h * w
        out = torch.empty(b, ncb * cb.g, h, w, dtype=torch.float32, device=indices.device)
    else:
        if indices.dim() != 3:
            raise RuntimeError("blc indices must be [b, l, ncb]")
        b, l, ncb = indices.shape
        out = torch.empty(b, l, ncb * cb.g, dtype=torch.float32, device=indices.device)
    flag = torch.zeros(1, dtype=torch.int32, device=indices.device) if check_range else None
    with torch.cuda.device(indices.device):
        fn = _lib.lib().gq_dequant_packed if packed else _lib.lib().gq_dequant
        check(fn(cb.handle, _ptr(indices), _ptr(out), b, l, ncb, layout, _ptr(flag), _stream()), "gq_dequant")
    if check_range and int(flag.item()):
        raise IndexError("index out of range in dequant")  # torch.index_select raises likewise
    return out


# ---- index stream utilities ----------------------------------------------------
def packed_dtype(n_samples: int) -> torch.dtype:
    """dtype of the packed index stream of an n_samples-entry codebook (gq_index_width)."""
    return torch.uint16 if int(_lib.lib().gq_index_width(int(n_samples))) == 2 else torch.int32


def _require_cuda_idx(indices: torch.Tensor) -> torch.Tensor:
    if indices.dtype != torch.int64:
        raise RuntimeError(f"indices must be int64, got {indices.dtype}")
    if not indices.is_cuda:
        raise NotImplementedError("indices must be a CUDA tensor (no CPU implementation exists)")
    return indices.contiguous()


def pack_indices(indices: torch.Tensor, n_samples: int, check_range: bool = False) -> torch.Tensor:
    """int64 indices (any shape) -> uint16 (n_samples <= 65536) / int32 stream of the same shape.
    Opt-in wire format: the reference carries int64 everywhere (gaussian.py:138,159)."""
    indices = _require_cuda_idx(indices)
    out = torch.empty(indices.shape, dtype=packed_dtype(n_samples), device=indices.device)
    flag = torch.zeros(1, dtype=torch.int32, device=indices.device) if check_range else None
    with torch.cuda.device(indices.device):
        check(_lib.lib().gq_pack_indices(_ptr(indices), indices.numel(), int(n_samples), _ptr(out), _ptr(flag),
                                         _stream()), "gq_pack_indices")
    if check_range and int(flag.item()):
        raise IndexError("index out of range in pack_indices")
    return out


def unpack_indices(packed: torch.Tensor, n_samples: int) -> torch.Tensor:
    """Inverse of pack_indices: -> int64, same shape."""
    if packed.dtype != packed_dtype(n_samples):
        raise RuntimeError(f"a {n_samples}-entry codebook packs to {packed_dtype(n_samples)}, got {packed.dtype}")
    if not packed.is_cuda:
        raise NotImplementedError("packed indices must be a CUDA tensor (no CPU implementation exists)")
    packed = packed.contiguous()
    out = torch.empty(packed.shape, dtype=torch.int64, device=packed.device)
    with torch.cuda.device(packed.device):
        check(_lib.lib().gq_unpack_indices(_ptr(packed), packed.numel(), int(n_samples), _ptr(out), _stream()),
              "gq_unpack_indices")
    return out


def index_histogram(indices: torch.Tensor, n_samples: int, hist: Optional[torch.Tensor] = None,
                    check_range: bool = False) -> torch.Tensor:
    """hist[i] += #(indices == i) (eval.py:127,152-154: all_hist += torch.histogram(indices, range(0, N))).
    hist: int64 [n_samples] on the same device, accumulated in place (created zeroed when None)."""
    indices = _require_cuda_idx(indices)
    if hist is None:
        hist = torch.zeros(int(n_samples), dtype=torch.int64, device=indices.device)
    if hist.dtype != torch.int64 or hist.numel() != int(n_samples) or not hist.is_contiguous() or hist.device != indices.device:
        raise RuntimeError("hist must be a contiguous int64 tensor of n_samples bins on the indices' device")
    flag = torch.zeros(1, dtype=torch.int32, device=indices.device) if check_range else None
    with torch.cuda.device(indices.device):
        check(_lib.lib().gq_index_histogram(_ptr(indices), indices.numel(), int(n_samples), _ptr(hist), _ptr(flag),
                                            _stream()), "gq_index_histogram")
    if check_range and int(flag.item()):
        raise IndexError("index out of range in index_histogram")
    return hist


def usage_entropy(hist: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """cal_ent (eval.py:137-141): (codebook usage in [0,1], entropy in bits) as 0-dim fp32 tensors."""
    if hist.dtype != torch.int64 or not hist.is_cuda:
        raise RuntimeError("hist must be the int64 CUDA tensor index_histogram accumulates")
    hist = hist.contiguous()
    out = torch.empty(2, dtype=torch.float32, device=hist.device)
    with torch.cuda.device(hist.device):
        check(_lib.lib().gq_usage_entropy(_ptr(hist), hist.numel(), _ptr(out), _stream()), "gq_usage_entropy")
    return out[0], out[1]


class _KlPenalty(torch.autograd.Function):
    """kl_fwd/kl_bwd pair.  Outputs (zhat, kl_loss, stats[8]); stats is non-differentiable."""

    @staticmethod
    def forward(ctx, z, eps, C, g, layout, lv_lo, lv_hi, target_bits, tol, lam, lam_min, lam_max):
        _require_cuda_f32("z", z)
        z = z.contiguous()
        if layout == GQ_LAYOUT_BCHW:
            b, c2, h, w = z.shape
            l = h * w
            zhat = torch.empty(b, C, h, w, dtype=torch.float32, device=z.device)
        else:
            b, l, c2 = z.shape
            zhat = torch.empty(b, l, C, dtype=torch.float32, device=z.device)
        if c2 != 2 * C:
            raise RuntimeError("latent must carry 2*C channels (mu | logvar)")
        if eps is not None:
            _require_cuda_f32("eps", eps)
            eps = eps.contiguous()
            if eps.shape != zhat.shape:
                raise RuntimeError("eps must have the shape of mu")
        ncb = C // g
        w_out = torch.empty(b * l * ncb, dtype=torch.float32, device=z.device)
        stats = torch.empty(8, dtype=torch.float32, device=z.device)
        L = _lib.lib()
        ws = torch.empty(int(L.gq_kl_workspace_bytes(b, l, C)), dtype=torch.uint8, device=z.device)
        with torch.cuda.device(z.device):
            check(L.gq_kl_fwd(_ptr(z), _ptr(eps), _ptr(zhat), _ptr(w_out), _ptr(stats), b, l, C, g, layout,
                              lv_lo, lv_hi, target_bits, tol, lam, lam_min, lam_max, _ptr(ws), ws.numel(),
                              _stream()), "gq_kl_fwd")
        ctx.save_for_backward(z, eps, w_out)
        ctx.cfg = (b, l, C, g, layout, lv_lo, lv_hi, lam)
        kl_loss = stats[0]
        ctx.mark_non_differentiable(stats)
        return zhat, kl_loss, stats

    @staticmethod
    def backward(ctx, grad_zhat, grad_kl, _grad_stats):
        z, eps, w_out = ctx.saved_tensors
        b, l, C, g, layout, lv_lo, lv_hi, lam = ctx.cfg
        grad_z = torch.empty_like(z)
        if grad_zhat is not None:
            grad_zhat = grad_zhat.contiguous().float()
        if grad_kl is not None:
            grad_kl = grad_kl.contiguous().float()
        with torch.cuda.device(z.device):
            check(_lib.lib().gq_kl_bwd(_ptr(z), _ptr(eps), _ptr(w_out), _ptr(grad_zhat), _ptr(grad_kl),
                                       _ptr(grad_z), b, l, C, g, layout, lv_lo, lv_hi, lam, _stream()),
                  "gq_kl_bwd")
        return (grad_z,) + (None,) * 11


def kl_penalty(z: torch.Tensor, eps: Optional[torch.Tensor], group: int, layout: int,
               logvar_range=(-30.0, 20.0), target_bits: float = 16.0, tolerance: float = 0.5,
               lam: float = 1.0, lam_min: float = 1.0, lam_max: float = 1.0):
    """Train-branch penalty (gaussian.py:76-98) -> (zhat, kl_loss, stats) with
    stats = [kl_loss, bits_mean, bits_min, bits_max, sum_kl2, sum_w_kl2, 0, 0]."""
    C = (z.shape[1] if layout == GQ_LAYOUT_BCHW else z.shape[2]) // 2
    if C % group:
        raise RuntimeError(f"channels {C} not divisible by group {group}")
    return _KlPenalty.apply(z, eps, C, int(group), int(layout), float(logvar_range[0]),
                            float(logvar_range[1]), float(target_bits), float(tolerance), float(lam),
                            float(lam_min), float(lam_max))
