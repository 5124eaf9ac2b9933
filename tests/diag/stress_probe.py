"""Stress latents (std = |N(0,1)|, reference test_extension.py:18-19): time of quantize_rows and a full
oracle check of the first rows; the rows beyond the split-image window take the filtered sliced exact scan."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, gq_b200, oracle
from gq_b200_host import host_codebook
dev = torch.device("cuda:0")
check = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for g, N, M in ((16, 65536, 65536), (8, 65536, 65536), (16, 4096, 65536), (32, 16384, 8192)):
    cbn, nlp = host_codebook(N, g)
    cb = gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
    gen = torch.Generator().manual_seed(g + N % 1000)
    mu = torch.randn(M, g, generator=gen)
    std = torch.exp(0.5 * (2.0 * torch.log(torch.abs(torch.randn(M, g, generator=gen)) + 1e-3)))
    lstd = std.log()
    a = (mu.to(dev), std.to(dev), lstd.to(dev))
    out = {}
    for mode in ("auto", "fast"):
        cb.set_mode(mode)
        for _ in range(3): idx, zh = gq_b200.quantize_rows(cb, *a)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx, zh = gq_b200.quantize_rows(cb, *a); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        st = torch.zeros(8, dtype=torch.int32, device=dev)
        idx, zh = gq_b200.quantize_rows(cb, *a, stats=st)
        n = min(check, M)
        o_idx, o_zh, _, _ = oracle.quantize(mu.numpy()[:n], std.numpy()[:n], lstd.numpy()[:n], cbn, nlp)
        out[mode] = {"ms": float(np.median(ts)), "rows_to_exact_scan": int(st[2]), "mismatches": int((idx.cpu().numpy()[:n] != o_idx).sum()),
                     "zhat_equal": bool(np.array_equal(zh.cpu().numpy()[:n], o_zh)), "checked": n}
    print(json.dumps({"g": g, "N": N, "M": M, **out}), flush=True)
