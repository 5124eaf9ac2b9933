#!/usr/bin/env python
"""BASELINE.json configs[4]-style sweep of the fused quantizer on one GPU: codebook size N x rows M
(x group g), CUDA-event time of `quantize_rows` (mu/std rows in HBM -> indices + codewords), rows/s,
algorithmic TFLOP/s (4 g M N flop) and the fraction of the measured tensor peak, with a
bit-exact spot check of 512 rows per point against the CPU oracle.  Prints one JSON object.

    python tests/diag/sweep.py [--quick]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import gq_b200  # noqa: E402
import oracle  # noqa: E402
from gq_b200_host import host_codebook  # noqa: E402


def rows(M, g, seed, dist="trained"):
    """The three named latent distributions of SURVEY 8(d)."""
    gen = torch.Generator().manual_seed(seed)
    if dist == "trained":       # what a converged model emits: ~16 bits per 16-dim group
        mu = torch.randn(M, g, generator=gen) * 0.866
        lv = -1.386 + 0.3 * torch.randn(M, g, generator=gen)
    elif dist == "randinit":    # what the random-init UNet of configs[1] emits
        mu = torch.randn(M, g, generator=gen) * 0.34
        lv = torch.randn(M, g, generator=gen) * 0.34
    else:                       # "stress": gq_cuda_extension/test/test_extension.py:18-19, std = |N(0,1)|
        mu = torch.randn(M, g, generator=gen)
        lv = 2.0 * torch.log(torch.abs(torch.randn(M, g, generator=gen)) + 1e-3)
    std = torch.exp(0.5 * lv)
    return mu, std, std.log()


def main():
    quick = "--quick" in sys.argv
    peak = 1648.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        peak = float(d.get("bf16_tflops", d.get("tensor_tflops", peak)))
    dev = torch.device("cuda:0")
    points = []
    grid = [(16, N, M, "trained") for N in (4096, 16384, 65536, 262144) for M in (1024, 16384, 65536, 262144, 1048576)]
    grid += [(4, 65536, 1048576, "trained"), (8, 65536, 524288, "trained"), (4, 65536, 65536, "trained"),
             (8, 65536, 65536, "trained")]                                     # configs[2]-style sub-codebooks
    grid += [(16, 65536, 65536, "randinit"), (16, 65536, 65536, "stress"), (16, 65536, 8192, "stress")]
    if quick:
        grid = [(16, 65536, 65536, "trained"), (4, 65536, 65536, "trained"), (16, 65536, 8192, "stress")]
    cbs = {}
    for g, N, M, dist in grid:
        if (N, g) not in cbs:
            cbn, nlp = host_codebook(N, g)
            cbs[(N, g)] = (cbn, nlp, gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0))
        cbn, nlp, cb = cbs[(N, g)]
        mu, std, lstd = rows(M, g, N % 1000 + M % 1000 + g, dist)
        mu_d, std_d, lstd_d = mu.to(dev), std.to(dev), lstd.to(dev)
        stats = torch.zeros(8, dtype=torch.int32, device=dev)
        for _ in range(3):
            idx, zhat = gq_b200.quantize_rows(cb, mu_d, std_d, lstd_d)
        torch.cuda.synchronize()
        iters = 5 if (M * N >= (1 << 36) or dist == "stress") else 20
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx, zhat = gq_b200.quantize_rows(cb, mu_d, std_d, lstd_d)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        gq_b200.quantize_rows(cb, mu_d, std_d, lstd_d, stats=stats)
        sel = np.random.RandomState(1).choice(M, min(512, M), replace=False)
        o_idx, _, _, _ = oracle.quantize(mu.numpy()[sel], std.numpy()[sel], lstd.numpy()[sel], cbn, nlp)
        mism = int((idx.cpu().numpy()[sel] != o_idx).sum())
        flop = 4.0 * g * M * N
        st = stats.cpu().tolist()
        points.append({"g": g, "N": N, "M": M, "latents": dist, "ms": ms, "rows_per_s": M / ms * 1e3, "tflops": flop / ms / 1e9,
                       "frac_of_tensor_peak": flop / ms / 1e9 / peak, "rows_checked": int(len(sel)), "index_mismatches": mism,
                       "rows_to_exact_scan": st[2], "rows_settled_in_exact_order": st[4]})
        print(points[-1], file=sys.stderr, flush=True)
    print(json.dumps({"tensor_peak_tflops": peak, "timing": "median CUDA-event time of quantize_rows (prep + search + rescoring + fallback), rows resident in HBM",
                      "points": points}))


if __name__ == "__main__":
    main()
