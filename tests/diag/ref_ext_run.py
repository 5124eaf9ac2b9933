#!/usr/bin/env python
"""Runs the REFERENCE's own CUDA extension (oracle/_ref/gq_cuda_ref.so, built from /root/reference by
oracle/build_ref.sh) in a process of its own: that library and our gq_cuda package both register the
operator `extension_cpp::gq`, so they cannot share an interpreter.  Test / diagnostic infrastructure.

    python tests/diag/ref_ext_run.py scores IN.npz OUT.npy      # out[M,N] of the reference kernel on (mu, std, noise, beta)
    python tests/diag/ref_ext_run.py time M N                   # JSON: reference gq_cuda + torch.argmax + index_select, ms
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "oracle", "_ref", "gq_cuda_ref.so")


def load():
    torch.ops.load_library(SO)
    return torch.ops.extension_cpp.gq


def main():
    gq = load()
    if sys.argv[1] == "scores":
        d = np.load(sys.argv[2])
        mu, std, noise = (torch.from_numpy(d[k]).cuda() for k in ("mu", "std", "noise"))
        M, dim = mu.shape
        N = noise.shape[0]
        out = torch.empty(M, N, device="cuda")
        gq(mu, std, noise, out, dim, M, N, float(d["beta"]))
        torch.cuda.synchronize()
        np.save(sys.argv[3], out.cpu().numpy())
        return
    M, N = int(sys.argv[2]), int(sys.argv[3])
    dim = 16
    gen = torch.Generator().manual_seed(0)
    noise = torch.randn(N, dim, generator=gen).cuda()
    mu = (torch.randn(M, dim, generator=gen) * 0.866).cuda()
    std = torch.exp(0.5 * (-1.386 + 0.3 * torch.randn(M, dim, generator=gen))).cuda()
    out = torch.empty(M, N, device="cuda")      # the reference's `self.perturbed` buffer (gaussian.py:126-128)

    def step():     # the cuda backend of the eval branch, gaussian.py:124-133
        gq(mu, std, noise, out, dim, M, N, 1.0)
        idx = torch.argmax(out, dim=1).clone()
        return torch.index_select(noise, 0, idx)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    e0.record()
    for _ in range(iters):
        gq(mu, std, noise, out, dim, M, N, 1.0)
    e1.record()
    torch.cuda.synchronize()
    ms_k = e0.elapsed_time(e1) / iters
    print(json.dumps({"rows": M, "n_samples": N, "ms_per_call": ms, "rows_per_s": M / ms * 1e3, "score_kernel_ms": ms_k,
                      "score_matrix_bytes": M * N * 4, "what": "reference gq_cuda_kernel + torch.argmax + index_select (gaussian.py:124-133), B200"}))


if __name__ == "__main__":
    main()
