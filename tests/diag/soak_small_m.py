"""Parity soak of the small-M plans (more than 8 storage segments per row -> rescore_many_kernel): every index
and every returned codeword of quantize_rows against the CPU oracle, over codebook sizes, groups, row counts
around the unit / tile boundaries and the three latent distributions of SURVEY 8(d)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, gq_b200, oracle
from gq_b200_host import host_codebook

def rows(M, g, seed, dist):
    gen = torch.Generator().manual_seed(seed)
    if dist == "trained":
        mu = torch.randn(M, g, generator=gen) * 0.866; lv = -1.386 + 0.3 * torch.randn(M, g, generator=gen)
    elif dist == "randinit":
        mu = torch.randn(M, g, generator=gen) * 0.34; lv = torch.randn(M, g, generator=gen) * 0.34
    else:
        mu = torch.randn(M, g, generator=gen); lv = 2.0 * torch.log(torch.abs(torch.randn(M, g, generator=gen)) + 1e-3)
    std = torch.exp(0.5 * lv)
    return mu, std, std.log()

dev = torch.device("cuda:0")
t0 = time.time()
cases, total_rows, total_bad, many = [], 0, 0, 0
dists = ("trained", "randinit", "stress")
k = 0
for N in (4096, 16384, 65536, 262144):
    for g in (16, 8, 4):
        cbn, nlp = host_codebook(N, g)
        cb = gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
        for M in (1, 5, 64, 255, 256, 257, 700, 1024, 1500, 3000):
            dist = dists[k % 3]; k += 1
            mu, std, lstd = rows(M, g, 1000 * k + M, dist)
            st = torch.zeros(8, dtype=torch.int32, device=dev)
            idx, zh = gq_b200.quantize_rows(cb, mu.to(dev), std.to(dev), lstd.to(dev), stats=st)
            o_idx, o_zh, _, _ = oracle.quantize(mu.numpy(), std.numpy(), lstd.numpy(), cbn, nlp)
            bad = int((idx.cpu().numpy() != o_idx).sum()) + (0 if np.array_equal(zh.cpu().numpy(), o_zh) else 1)
            s = st.cpu().tolist()
            cases.append({"N": N, "g": g, "M": M, "latents": dist, "mismatches": bad, "rows_to_exact_scan": s[2], "rows_settled_exact_order": s[4]})
            total_rows += M; total_bad += bad
            print(cases[-1], file=sys.stderr, flush=True)
print(json.dumps({"rows": total_rows, "cases": len(cases), "mismatches": total_bad, "seconds": round(time.time() - t0, 1), "per_case": cases}))
