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
    """Median CUDA-event time of one launch.  The L2 is flushed before every timed launch, and a
    device-side spin keeps the GPU busy while the host enqueues event + launch + event, so the
    interval holds the kernel and not the host's launch path."""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()          # 256 MB write: evicts L2 between timed launches
        torch.cuda._sleep(400000)  # ~0.2 ms of device time: the host gets ahead of the queue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def raw_kl(L, z, eps, g, layout, b, l, C):
    """Pre-allocated buffers + direct C-ABI calls (no autograd / allocator time in the interval)."""
    from gq_b200.ops import _ptr, _stream, check
    ncb = C // g
    zhat = torch.empty_like(eps)
    w_out = torch.empty(b * l * ncb, dtype=torch.float32, device=z.device)
    stats = torch.empty(8, dtype=torch.float32, device=z.device)
    ws = torch.empty(int(L.gq_kl_workspace_bytes(b, l, C)), dtype=torch.uint8, device=z.device)
    gz = torch.randn_like(eps)
    gk = torch.ones(1, device=z.device)
    grad_z = torch.empty_like(z)
    s = _stream()

    def fwd():
        check(L.gq_kl_fwd(_ptr(z), _ptr(eps), _ptr(zhat), _ptr(w_out), _ptr(stats), b, l, C, g, layout, -30.0, 20.0,
                          16.0, 0.5, 1.0, 1.0, 1.0, _ptr(ws), ws.numel(), s), "gq_kl_fwd")

    def bwd():
        check(L.gq_kl_bwd(_ptr(z), _ptr(eps), _ptr(w_out), _ptr(gz), _ptr(gk), _ptr(grad_z), b, l, C, g, layout,
                          -30.0, 20.0, 1.0, s), "gq_kl_bwd")
    return fwd, bwd


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
    from gq_b200 import _lib
    L = _lib.lib()

    def rec(name, ms, nbytes):
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}

    z = z.detach()
    nrow = b * l * (C // g)
    fwd_bytes = z.numel() * 4 + eps.numel() * 8 + nrow * 4          # read z, eps; write zhat, w
    bwd_bytes = z.numel() * 8 + eps.numel() * 8 + nrow * 4          # read z, eps, grad_zhat, w; write grad_z
    fwd, bwd = raw_kl(L, z, eps, g, gq_b200.ops.GQ_LAYOUT_BLC, b, l, C)
    rec("kl_fwd_blc", timed(fwd, flush=flush), fwd_bytes)
    rec("kl_bwd_blc", timed(bwd, flush=flush), bwd_bytes)
    # bchw variant (sd3unet configs), batch 256
    zb = torch.cat([torch.randn(256, C, 32, 32) * 0.866, -1.386 + 0.3 * torch.randn(256, C, 32, 32)], 1).to(dev)
    epsb = torch.randn(256, C, 32, 32, device=dev)
    fwd, bwd = raw_kl(L, zb, epsb, g, gq_b200.ops.GQ_LAYOUT_BCHW, 256, 1024, C)
    rec("kl_fwd_bchw", timed(fwd, flush=flush), fwd_bytes)
    rec("kl_bwd_bchw", timed(bwd, flush=flush), bwd_bytes)
    # yardstick: a plain device copy moving the same number of bytes as kl_fwd (half read, half written)
    src = torch.empty(fwd_bytes // 8, dtype=torch.float32, device=dev)
    dst = torch.empty_like(src)
    rec("copy_same_bytes_as_kl_fwd", timed(lambda: dst.copy_(src), flush=flush), fwd_bytes)
    # --- dequant, batch 256 @ 32x32 tokens, group 16 (8 B index + 4 g B output per row)
    cbn, nlp = host_codebook(65536, 16)
    cb = gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
    for name, shape, layout in (("dequant_bchw", (256, 1, 32, 32), gq_b200.ops.GQ_LAYOUT_BCHW),
                                ("dequant_blc", (256, 1024, 1), gq_b200.ops.GQ_LAYOUT_BLC)):
        from gq_b200.ops import _ptr, _stream, check
        idx = torch.randint(0, 65536, shape, device=dev)
        o16 = torch.empty(idx.numel() * 16, dtype=torch.float32, device=dev)
        bb, ll = 256, 1024
        ms = timed(lambda: check(L.gq_dequant(cb.handle, _ptr(idx), _ptr(o16), bb, ll, 1, layout, None, _stream()), "dq"),
                   flush=flush)
        rec(name, ms, idx.numel() * 8 + idx.numel() * 16 * 4)
    # --- index stream: histogram (8 B read per index), pack / unpack (8 B + 2 B per index), 16 M indices
    n_idx = 1 << 24
    big = torch.randint(0, 65536, (n_idx,), device=dev)
    hist = torch.zeros(65536, dtype=torch.int64, device=dev)
    rec("index_histogram", timed(lambda: gq_b200.index_histogram(big, 65536, hist), iters=20, flush=flush), n_idx * 8)
    rec("pack_indices_u16", timed(lambda: gq_b200.pack_indices(big, 65536), iters=20, flush=flush), n_idx * 10)
    pk = gq_b200.pack_indices(big, 65536)
    rec("unpack_indices_u16", timed(lambda: gq_b200.unpack_indices(pk, 65536), iters=20, flush=flush), n_idx * 10)
    del big, pk
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
