#!/usr/bin/env python
"""HBM-bound kernels of the path (KL penalty fwd/bwd, dequant gather, legacy score matrix, fused
front-end) timed alone with CUDA events; achieved GB/s on ALGORITHMIC bytes against the measured
copy bandwidth (MEASURED_PEAKS.json).  Prints one JSON object; run under gpurun."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gq_b200  # noqa: E402
from gq_b200_host import host_codebook  # noqa: E402


def timed(fn, iters=50, flush=None):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()          # 256 MB write: evicts L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    peak = 6559.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"hbm_peak_gbs": peak, "l2_flush_between_launches": True}
    # --- KL penalty, BASELINE configs[3]: z [256, 1024, 32] blc, group 16
    b, l, C, g = 256, 1024, 16, 16
    z = torch.cat([torch.randn(b, l, C) * 0.866, -1.386 + 0.3 * torch.randn(b, l, C)], 2).to(dev).requires_grad_(True)
    eps = torch.randn(b, l, C, device=dev)
    res = {}

    def fwd():
        res["o"] = gq_b200.kl_penalty(z, eps, g, gq_b200.ops.GQ_LAYOUT_BLC, (-30.0, 20.0), 16.0, 0.5, 1.0, 1.0, 1.0)
    ms = timed(fwd, flush=flush)
    nbytes = z.numel() * 4 + eps.numel() * 4 + eps.numel() * 4 + (b * l * (C // g)) * 4   # read z, eps; write zhat, w
    out["kl_fwd_blc"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    zhat, kl, stats = res["o"]
    gz = torch.randn_like(zhat)
    gk = torch.ones((), device=dev)

    def bwd():
        torch.autograd.grad([zhat, kl], [z], [gz, gk], retain_graph=True)
    ms = timed(bwd, flush=flush)
    nbytes = z.numel() * 4 * 2 + eps.numel() * 4 * 2 + (b * l * (C // g)) * 4   # read z, eps, grad_zhat, w; write grad_z
    out["kl_bwd_blc"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    # bchw variant (sd3unet configs), batch 256
    zb = torch.cat([torch.randn(256, C, 32, 32) * 0.866, -1.386 + 0.3 * torch.randn(256, C, 32, 32)], 1).to(dev)
    epsb = torch.randn(256, C, 32, 32, device=dev)
    ms = timed(lambda: gq_b200.kl_penalty(zb, epsb, g, gq_b200.ops.GQ_LAYOUT_BCHW, (-30.0, 20.0), 16.0, 0.5, 1.0, 1.0, 1.0), flush=flush)
    nbytes = zb.numel() * 4 + epsb.numel() * 8 + 256 * 1024 * 4
    out["kl_fwd_bchw"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    # --- dequant, batch 256 @ 32x32 tokens, group 16 (8 B index + 4 g B output per row)
    cbn, nlp = host_codebook(65536, 16)
    cb = gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
    for name, shape, layout in (("dequant_bchw", (256, 1, 32, 32), gq_b200.ops.GQ_LAYOUT_BCHW),
                                ("dequant_blc", (256, 1024, 1), gq_b200.ops.GQ_LAYOUT_BLC)):
        idx = torch.randint(0, 65536, shape, device=dev)
        ms = timed(lambda: gq_b200.dequant(cb, idx, layout), flush=flush)
        nbytes = idx.numel() * 8 + idx.numel() * 16 * 4
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    # --- legacy score matrix: M=8192 rows x 65536 (2.1 GB fp32 out, write-bound)
    M = 8192
    mu = torch.randn(M, 16, device=dev)
    sd = torch.rand(M, 16, device=dev) + 0.3
    o = torch.empty(M, 65536, device=dev)
    import gq_cuda
    ms = timed(lambda: gq_cuda.ops.gq_cuda(mu, sd, cb_t(cbn, dev), o, 16, M, 65536, 1.0), iters=5)
    nbytes = o.numel() * 4
    out["legacy_scores"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak,
                            "note": "16 IEEE divides per pair keep it CUDA-core bound, as the reference kernel is"}
    print(json.dumps(out))


_CB = {}


def cb_t(cbn, dev):
    if "t" not in _CB:
        _CB["t"] = torch.from_numpy(cbn).to(dev)
    return _CB["t"]


if __name__ == "__main__":
    main()
