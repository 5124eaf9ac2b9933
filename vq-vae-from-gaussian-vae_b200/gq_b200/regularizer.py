"""Drop-in `GaussianQuantRegularizer` for B200.

Mirrors pit/quantization/gaussian.py:22-178 of the reference -- same constructor arguments,
same `forward(z) -> (zhat, info)` / `dequant(indices)` protocol, same `info` keys, same
attribute names (`prior_samples`, `normal_log_prob`, `lam`, `lam_min`, `lam_max`) -- so that
`target: pit.quantization.gaussian.GaussianQuantRegularizer` configs keep working when this
class is bound at that path.  The hot parts run in libgq_b200:

  eval  : score + argmax + gather   -> gq_quantize   (tcgen05 search + exact re-score)
  train : KL penalty + reparam      -> gq_kl_fwd/bwd (single pass)
  dequant: gather + permutes        -> gq_dequant

There is no CPU or pure-PyTorch fallback: CPU tensors raise, a missing library fails at
import.  `backend` is accepted for config compatibility ("torch" and "cuda" both select the
CUDA path; anything else raises ValueError at forward time as the reference does).
What deliberately stays in PyTorch: the codebook generator `prior_samples()` (host, once),
`torch.randn_like` (RNG parity with the reference), and the exp/log/clamp prologue on rows
(so std and log(std) are the framework's own values, as in the reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .ops import GQ_LAYOUT_BCHW, GQ_LAYOUT_BLC, Codebook

__all__ = ["prior_samples", "GaussianQuantRegularizer", "GaussianQuantRegularizer2"]


from gq_b200_host import normal_log_prob as _normal_log_prob, prior_samples  # noqa: E402,F401


def _reference_noise(b: int, l: int, c: int, bchw: bool, hw, device) -> torch.Tensor:
    """torch.randn_like(mu) exactly as the reference issues it (gaussian.py:84,121), returned in
    zhat's layout.  mu is a chunk VIEW of the blc latent; in bchw format that latent is itself a
    rearrange view of z, so mu has strides (2c*l, 1, l) and randn_like fills memory in [b,c,l]
    order -- i.e. the noise comes out as a contiguous [b,c,h,w] tensor.  Reproducing the same
    call on an empty tensor with those strides keeps the RNG stream bit-identical."""
    if bchw:
        proto = torch.empty_strided((b, l, c), (2 * c * l, 1, l), dtype=torch.float32, device=device)
        eps = torch.randn_like(proto)               # strides (c*l, 1, l)
        return eps.transpose(1, 2).reshape(b, c, *hw)   # free view: contiguous [b,c,h,w]
    proto = torch.empty_strided((b, l, c), (l * 2 * c, 2 * c, 1), dtype=torch.float32, device=device)
    return torch.randn_like(proto)


class _GQBase(nn.Module):
    """State shared by both regularizer variants: codebook buffers + prepared device handle."""

    def _init_codebook(self, n_samples: int, dim: int, seed: int, beta: float) -> None:
        self.register_buffer("prior_samples", prior_samples(n_samples, dim, seed).float(), persistent=False)
        self.register_buffer("normal_log_prob", _normal_log_prob(self.prior_samples).float(), persistent=False)
        self._cb: Optional[Codebook] = None
        self._cb_key = None

    # The prepared device codebook is a native handle (ctypes pointer): it cannot be pickled or deep-copied, and two
    # modules must never own one handle (its finaliser frees device memory).  Copies therefore drop it and
    # rebuild it lazily on their first eval forward / dequant -- the reference module can be deep-copied (EMA
    # copies) and pickled (torch.save(model)) at any time, and so can this one.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_cb"] = None
        state["_cb_key"] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_cb", "_cb_key") else copy.deepcopy(v, memo)
        return new

    #: opt-in range check of index tensors in dequant / pack_indices / accumulate_usage: reading the flag back is a
    #: blocking device-to-host sync, which the reference's index_select never does (gaussian.py:171-173; on a GPU it
    #: raises a device-side assert asynchronously).  Off: out-of-range ids read codeword 0.  Also GQ_B200_CHECK_RANGE=1.
    check_range = False

    def _check_range(self) -> bool:
        import os
        return bool(self.check_range) or os.environ.get("GQ_B200_CHECK_RANGE", "0") == "1"

    def _codebook(self) -> Codebook:
        ps = self.prior_samples
        if not ps.is_cuda:
            raise NotImplementedError(
                "GaussianQuantRegularizer (B200) has no CPU path: move the module to a CUDA device")
        key = (ps.data_ptr(), ps.device, float(self.beta), self.normal_log_prob.data_ptr())
        if self._cb is None or self._cb_key != key:
            self._cb = Codebook(ps, self.normal_log_prob, float(self.beta))
            self._cb_key = key
            self._cb.set_mode(getattr(self, "_search_mode", "auto"))
            self._cb.set_usage_histogram(getattr(self, "_usage_hist", None))
        return self._cb

    def track_usage(self, on: bool = True):
        """Not in the reference module: accumulate the codebook-usage histogram of every eval forward inside
        the quantizer itself (eval.py:127,152-154 keeps it on the host).  Returns the int64 [n_samples]
        histogram tensor (None when switched off); `codebook_usage(hist)` turns it into (usage, entropy)."""
        if on:
            if getattr(self, "_usage_hist", None) is None:
                self._usage_hist = torch.zeros(self.n_samples, dtype=torch.int64, device=self.prior_samples.device)
        else:
            self._usage_hist = None
        if getattr(self, "_cb", None) is not None:
            self._cb.set_usage_histogram(self._usage_hist)
        return self._usage_hist

    def set_search_mode(self, mode: str) -> None:
        """Not in the reference module: search precision for groups 8 / 16 ("auto", "fast", "robust";
        Codebook.set_mode).  Results are identical in every mode; only the speed on low-variance rows differs."""
        self._search_mode = mode
        if getattr(self, "_cb", None) is not None:
            self._cb.set_mode(mode)


class GaussianQuantRegularizer(_GQBase):
    # train() mode: sample as Gaussian VAE, log contains loss function
    # eval() mode: sample as VQ-VAE, log contains indices
    def __init__(self, format, n_samples, group=1, logvar_range=[-30.0, 20.0],  # common parameters
                 tolerance=0.5, lam_factor=1.01,  # train parameters
                 seed=42, beta=1.0, backend="torch",  # inference parameters
                 ):
        super().__init__()
        self.format = format
        assert self.format in ["bchw", "blc"]
        self.group = group
        self.logvar_range = logvar_range
        self.n_samples = n_samples
        self.log_n_samples = int(math.log(self.n_samples, 2))

        # training parameter setup (gaussian.py:40-45)
        self.lam_factor = lam_factor
        self.lam = 1.0
        self.lam_min = 1.0
        self.lam_max = 1.0
        self.lam_range = (1e-3, 1e3)
        self.tolerance = tolerance

        # inference parameter setup (gaussian.py:48-59)
        self.beta = beta
        self.seed = seed
        self._init_codebook(self.n_samples, self.group, self.seed, self.beta)
        self.backend = backend
        self.last_stats: Optional[torch.Tensor] = None

    def _update_lambdas(self, kl2_mean: float, kl2_min: float, kl2_max: float) -> None:
        """The stateful host-side controller, gaussian.py:100-115, fed with python floats."""
        if kl2_mean > self.log_n_samples:
            self.lam = self.lam * self.lam_factor
        else:
            self.lam = self.lam / self.lam_factor
        if kl2_max > self.log_n_samples + self.tolerance:
            self.lam_max = self.lam_max * self.lam_factor
        else:
            self.lam_max = self.lam_max / self.lam_factor
        self.lam_max = max(min(self.lam_max, self.lam_range[1]), 1.0)
        if kl2_min < self.log_n_samples - self.tolerance:
            self.lam_min = self.lam_min / self.lam_factor
        else:
            self.lam_min = self.lam_min * self.lam_factor
        self.lam_min = max(min(self.lam_min, 1.0), self.lam_range[0])

    # -- train branch ------------------------------------------------------------------------
    def _forward_train(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        layout = GQ_LAYOUT_BCHW if self.format == "bchw" else GQ_LAYOUT_BLC
        if layout == GQ_LAYOUT_BCHW:
            b, c2, h, w = z.shape
            l, c = h * w, c2 // 2
        else:
            b, l, c2 = z.shape
            c = c2 // 2
        eps = _reference_noise(b, l, c, layout == GQ_LAYOUT_BCHW, (h, w) if layout == GQ_LAYOUT_BCHW else None, z.device)
        zhat, kl_loss, stats = ops.kl_penalty(
            z, eps, self.group, layout, self.logvar_range, float(self.log_n_samples), self.tolerance,
            self.lam, self.lam_min, self.lam_max)
        # lambda controller (gaussian.py:100-115): one device->host copy instead of three syncs
        _, kl2_mean, kl2_min, kl2_max = stats[:4].tolist()
        self._update_lambdas(kl2_mean, kl2_min, kl2_max)
        info = {"kl_loss": kl_loss, "bits-mean": stats[1], "bits-min": stats[2], "bits-max": stats[3],
                "lam": torch.zeros_like(stats[0]) + self.lam}
        return zhat, info

    # -- eval branch -------------------------------------------------------------------------
    def _forward_eval(self, z: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        if self.backend not in ("cuda", "torch"):
            raise ValueError
        layout = GQ_LAYOUT_BCHW if self.format == "bchw" else GQ_LAYOUT_BLC
        if layout == GQ_LAYOUT_BCHW:
            b, c2, h, w = z.shape
            l, c = h * w, c2 // 2
        else:
            b, l, c2 = z.shape
            c = c2 // 2
        eps = _reference_noise(b, l, c, layout == GQ_LAYOUT_BCHW, (h, w) if layout == GQ_LAYOUT_BCHW else None, z.device)
        zq, indices, znq = ops.quantize_latent(self._codebook(), z, layout, eps, self.logvar_range,
                                               stats=self.last_stats)
        return zq, {"indices": indices, "zhat_noquant": znq}

    def forward(self, z):
        z = z.float()
        if not z.is_cuda:
            raise NotImplementedError(
                "GaussianQuantRegularizer (B200) has no CPU path: pass a CUDA tensor")
        if self.training:
            return self._forward_train(z)
        return self._forward_eval(z)

    def graphed_eval(self, z_example: torch.Tensor, warmup: int = 3):
        """Not in the reference module: capture ONE eval forward at `z_example`'s shape in a CUDA graph and
        return `run(z) -> (zhat, info)` that replays it.  A single-image call is launch-bound (seven dependent
        launches of ours plus torch's randn); the replay is one launch.  Measured on B200 (tests/diag/
        graph_probe.py): one 256-px image 92 -> 77 us isolated latency, 67 -> 57 us per call back to back;
        at batch 64 the eager call is faster (339 vs 361 us: the kernels dominate and the replay adds the copy
        into the static input), so this is for small batches.  `run` copies `z` (any float dtype, host or device) into the graph's static input, so it is
        `forward(z.float())` of the eval branch; the returned tensors are the graph's static outputs and are
        overwritten by the next `run`.  The noise of `zhat_noquant` comes from the default CUDA generator
        exactly as in eager mode (torch registers the generator with the graph and advances its offset on
        every replay).  The search mode in force at capture time is frozen into the graph (results do not
        depend on it)."""
        if self.training:
            raise RuntimeError("graphed_eval captures the eval branch: call .eval() first")
        if not z_example.is_cuda:
            raise NotImplementedError(
                "GaussianQuantRegularizer (B200) has no CPU path: pass a CUDA tensor")
        static_z = z_example.detach().float().clone()
        cur = torch.cuda.current_stream(static_z.device)
        side = torch.cuda.Stream(device=static_z.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):      # builds the codebook image and the workspace, settles the auto mode
                self._forward_eval(static_z)
        cur.wait_stream(side)
        keep = (self._codebook(), self._codebook()._ws)   # the graph holds raw pointers into these
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            zq, info = self._forward_eval(static_z)

        def run(z: torch.Tensor):
            static_z.copy_(z, non_blocking=True)
            graph.replay()
            return zq, info

        run.graph = graph
        run.static_input = static_z
        run._keep = keep
        return run

    def dequant(self, indices):
        """gaussian.py:162-178.  Also accepts the packed stream of `pack_indices` (uint16 / int32)."""
        layout = GQ_LAYOUT_BCHW if self.format == "bchw" else GQ_LAYOUT_BLC
        return ops.dequant(self._codebook(), indices, layout, check_range=self._check_range())

    # ---- not in the reference module: the index stream after the quantizer ----
    def pack_indices(self, indices):
        """int64 indices -> uint16 (n_samples <= 65536) / int32, same shape; `dequant` reads either."""
        return ops.pack_indices(indices, self.n_samples, check_range=self._check_range())

    def unpack_indices(self, packed):
        return ops.unpack_indices(packed, self.n_samples)

    def accumulate_usage(self, indices, hist=None):
        """all_hist += histogram(indices) over the codebook (eval.py:127,152-154); returns the int64 histogram."""
        return ops.index_histogram(indices, self.n_samples, hist, check_range=self._check_range())

    @staticmethod
    def codebook_usage(hist):
        """cal_ent (eval.py:137-141) -> (usage, entropy_bits) 0-dim tensors."""
        return ops.usage_entropy(hist)


class GaussianQuantRegularizer2(_GQBase):
    """Variant with contiguous grouping along `dim_idx` that runs the Gaussian and the VQ branch
    on every forward (gaussian.py:181-362), including its quirks: mean-reduced KL (:240-241),
    lam_range (1e-7, 1e7) (:196), and lam_max that never decreases because :251 is a no-op."""

    def __init__(self, dim, codebook_size, dim_idx=1, logvar_range=[-30.0, 20.0],
                 tolerance=0.5, lam_factor=1.01, seed=42, beta=1.0, use_ste=True, backend="torch"):
        super().__init__()
        self.dim_idx = dim_idx
        self.dim = dim
        self.logvar_range = logvar_range
        self.n_samples = codebook_size
        self.log_n_samples = int(math.log(self.n_samples, 2))
        self.lam_factor = lam_factor
        self.lam = 1.0
        self.lam_min = 1.0
        self.lam_max = 1.0
        self.lam_range = (1e-7, 1e7)
        self.tolerance = tolerance
        self.beta = beta
        self.seed = seed
        self._init_codebook(self.n_samples, self.dim, self.seed, self.beta)
        self.use_ste = use_ste
        self.backend = backend

    def quant_gaussian(self, z):
        z = torch.movedim(z.float(), self.dim_idx, -1)   # the kernels are fp32 (as GQ v1's z.float()); outputs are fp32
        assert z.shape[-1] % (self.dim * 2) == 0
        z_shape = z.shape
        z2 = z.reshape(-1, z_shape[-1])
        codebook_num = z2.shape[-1] // (self.dim * 2)
        C = codebook_num * self.dim
        rows = z2.shape[0]
        # contiguous grouping == the BLC kernel with ncb=1 on a [rows*codebook_num, 1, 2*dim]
        # view whose (mu | logvar) halves are gathered per sub-codebook
        mu, logvar = z2.chunk(2, -1)
        eps = torch.randn_like(mu)
        zk = torch.cat([mu.reshape(rows, codebook_num, self.dim), logvar.reshape(rows, codebook_num, self.dim)], -1)
        zk = zk.reshape(1, rows * codebook_num, 2 * self.dim)
        # mean-reduced loss: kl_loss = lam * mean(w*kl2) == (lam * sum(w*kl2) / 1) / count
        zhat_k, kl_sum, stats = ops.kl_penalty(
            zk, eps.reshape(1, rows * codebook_num, self.dim), self.dim, GQ_LAYOUT_BLC, self.logvar_range,
            float(self.log_n_samples), self.tolerance, self.lam, self.lam_min, self.lam_max)
        kl_loss = kl_sum / float(rows * codebook_num)
        _, kl2_mean, kl2_min, kl2_max = stats[:4].tolist()
        if kl2_mean > self.log_n_samples:
            self.lam = self.lam * self.lam_factor
        else:
            self.lam = self.lam / self.lam_factor
        if kl2_max > self.log_n_samples + self.tolerance:
            self.lam_max = self.lam_max * self.lam_factor
        # else: reference line 251 is an expression statement without assignment -> no change
        self.lam_max = max(min(self.lam_max, self.lam_range[1]), 1.0)
        if kl2_min < self.log_n_samples - self.tolerance:
            self.lam_min = self.lam_min / self.lam_factor
        else:
            self.lam_min = self.lam_min * self.lam_factor
        self.lam_min = max(min(self.lam_min, 1.0), self.lam_range[0])

        zhat = zhat_k.reshape(rows, C).reshape(*z_shape[:-1], -1)
        zhat = torch.movedim(zhat, -1, self.dim_idx)
        lv = torch.clamp(logvar, self.logvar_range[0], self.logvar_range[1])
        std = torch.exp(0.5 * lv)
        mu_o = torch.movedim(mu.reshape(*z_shape[:-1], -1), -1, self.dim_idx)
        std_o = torch.movedim(std.reshape(*z_shape[:-1], -1), -1, self.dim_idx)
        info = {"kl_loss": kl_loss, "bits-mean": stats[1], "bits-min": stats[2], "bits-max": stats[3],
                "lam-min": self.lam_min, "lam-max": self.lam_max, "lam": self.lam, "mu": mu_o, "std": std_o,
                "zhat_noquant": zhat}
        return zhat, info

    def quant_vq(self, z):
        if self.backend not in ("cuda", "torch"):
            raise ValueError
        z = torch.movedim(z.float(), self.dim_idx, -1)
        assert z.shape[-1] % (self.dim * 2) == 0
        z_shape = z.shape
        z2 = z.reshape(-1, z_shape[-1])
        codebook_num = z2.shape[-1] // (self.dim * 2)
        mu, logvar = z2.chunk(2, -1)
        logvar = torch.clamp(logvar, self.logvar_range[0], self.logvar_range[1])
        std = torch.exp(0.5 * logvar)
        mu = mu.reshape(-1, self.dim)
        std = std.reshape(-1, self.dim)
        idx, zq = ops.quantize_rows(self._codebook(), mu, std, std.log())
        zhat = zq.reshape(-1, codebook_num * self.dim).reshape(*z_shape[:-1], -1)
        zhat = torch.movedim(zhat, -1, self.dim_idx)
        indices = idx.reshape(-1, codebook_num).reshape(*z_shape[:-1], -1)
        indices = torch.movedim(indices, -1, self.dim_idx)
        return zhat, {"indices": indices, "zhat_quant": zhat}

    def forward(self, z):
        if not z.is_cuda:
            raise NotImplementedError(
                "GaussianQuantRegularizer2 (B200) has no CPU path: pass a CUDA tensor")
        zhat_g, info_g = self.quant_gaussian(z)
        with torch.no_grad():
            zhat_v, info_v = self.quant_vq(z)
        if self.use_ste:
            zhat = zhat_g - zhat_g.detach() + zhat_v
        else:
            zhat = zhat_g if self.training else zhat_v
        return zhat, info_g | info_v

    def dequant(self, indices):
        indices = torch.movedim(indices, self.dim_idx, -1)
        i_shape = indices.shape
        codebook_num = i_shape[-1]
        flat = indices.reshape(1, -1, 1).contiguous()
        zhat = ops.dequant(self._codebook(), flat, GQ_LAYOUT_BLC, check_range=self._check_range())
        zhat = zhat.reshape(-1, codebook_num * self.dim).reshape(*i_shape[:-1], -1)
        return torch.movedim(zhat, -1, self.dim_idx)
