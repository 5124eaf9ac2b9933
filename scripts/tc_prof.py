#!/usr/bin/env python
"""Timeline of the search kernel's warps.  Builds a -DGQ_TC_TRACE copy of the library (clock64 stamps at
the barrier wait / TMEM load / end of every epilogue iteration and around the waits and the issue of
the MMA issuers; CTA 0, iterations 300..323), runs BASELINE configs[1] rows through it and prints the
merged timeline.  Diagnostic only: the instrumented library is never shipped.

    python scripts/tc_prof.py build [extra nvcc flags]   # here, writes scripts/_build/libgq_b200_prof.so
    python scripts/tc_prof.py                            # on the GPU box
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200")
OUT = os.path.join(ROOT, "scripts", "_build", "libgq_b200_prof.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    src = [os.path.join(PKG, "csrc", f) for f in ("gq_api.cu", "gq_kernels.cu", "gq_kl.cu", "gq_tc.cu", "gq_tokens.cu")]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DGQ_TC_TRACE",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + src + sys.argv[2:]
    subprocess.check_call(cmd)
    print(OUT)


def main():
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    os.environ["GQ_B200_LIB"] = OUT
    import gq_b200
    from gq_b200 import _lib
    from gq_b200_host import host_codebook
    L = _lib.lib()
    dev = torch.device("cuda:0")
    N, g, M = 65536, 16, int(os.environ.get("GQ_PROF_ROWS", "65536"))
    first_it = int(os.environ.get("GQ_PROF_FIRST", "300"))     # must match -DGQ_TRACE_FROM of the build
    cbn, nlp = host_codebook(N, g)
    cb = gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
    gen = torch.Generator().manual_seed(1)
    mu = (torch.randn(M, g, generator=gen) * 0.866).to(dev)
    std = torch.exp(0.5 * (-1.386 + 0.3 * torch.randn(M, g, generator=gen))).to(dev)
    for _ in range(3):
        gq_b200.quantize_rows(cb, mu, std)
    torch.cuda.synchronize()
    n = 24
    buf = (ctypes.c_uint * (32 * n * 4))()
    L.gq_debug_tc_trace.argtypes = [ctypes.c_void_p]
    L.gq_debug_tc_trace(buf)
    t = np.array(list(buf), dtype=np.int64).reshape(32, n, 4)
    roles = {}
    for w in range(32):
        if t[w].any():
            roles[w] = w
    ts = os.environ.get("GQ_PROF_TS") == "1"      # the TS-form kernel: epilogue 0..11 (stage = w / 4), producer 12, issuers 13..16
    issuers = [13, 14, 15, 16] if ts else [17, 18]
    producer = 12 if ts else 16
    shown = (12, 13, 15, 0, 4, 8) if ts else (16, 17, 18, 0, 8, 4)
    first = 0 if first_it == 0 else min(t[w, 0, 0] for w in roles)   # small calls: cycles since the warp entered the kernel
    print("cycles relative to the first stamp.  I = issuer (top, B tile ready + stage drained, MMAs issued, committed + warp reconverged); "
          "E = epilogue warp (top, accumulator ready, data in registers, end of iteration)")
    ev = []
    for w in sorted(roles):
        name = f"I w{w}" if w in issuers else (f"P w{w}" if w == producer else f"E w{w}")
        if w not in shown:
            continue   # row block 0: its issuers (stage A / stage B) and two of its epilogue warps
        for i in range(14 if first_it == 0 else 10):
            if t[w, i, 0] == 0:
                continue
            ev.append((t[w, i, 0] - first, f"{name} it{first_it + i}: " + " ".join(f"{int(x - first):6d}" for x in t[w, i])))
    for _, line in sorted(ev):
        print("  " + line)
    for w in sorted(roles):
        rows = t[w][t[w, :, 0] > 0]          # (the TS-form kernel's issuers stamp every other iteration)
        if len(rows) < 2:
            continue
        d = np.diff(rows[:, 0])
        seg = [np.mean(rows[:, k + 1] - rows[:, k]) for k in range(3)]
        print(f"warp {w}: mean iteration {d.mean():.0f} cycles; segments {seg[0]:.0f} / {seg[1]:.0f} / {seg[2]:.0f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        main()
