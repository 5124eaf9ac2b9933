"""Stress latents at the reference extension-test shape: how does quantize_rows behave as the split-image window cap
is raised (libraries built with -DGQ_SPLIT_CAP_MUL=...)?  Usage: GQ_B200_LIB=... python tests/diag/stress_cap_probe.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, gq_b200, oracle
from gq_b200_host import host_codebook
dev = torch.device("cuda:0")
g, N, M = 16, 65536, 131072
cbn, nlp = host_codebook(N, g)
cb = gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
gen = torch.Generator().manual_seed(3)
mu = torch.randn(M, g, generator=gen)
std = torch.abs(torch.randn(M, g, generator=gen)).clamp_min(1e-12)
lstd = std.log()
a = (mu.to(dev), std.to(dev), lstd.to(dev))
for _ in range(4): gq_b200.quantize_rows(cb, *a)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx, zh = gq_b200.quantize_rows(cb, *a); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
st = torch.zeros(8, dtype=torch.int32, device=dev)
idx, zh = gq_b200.quantize_rows(cb, *a, stats=st)
n = 32768
o_idx, o_zh, _, _ = oracle.quantize(mu.numpy()[:n], std.numpy()[:n], lstd.numpy()[:n], cbn, nlp)
print(json.dumps({"lib": os.environ.get("GQ_B200_LIB", "default"), "ms": float(np.median(ts)), "rows_per_s": M / float(np.median(ts)) * 1e3,
                  "stats[rows,prefilter_cw,fallback,overflow,exact]": st.cpu().tolist()[:5],
                  "mismatches": int((idx.cpu().numpy()[:n] != o_idx).sum()), "checked": n}), flush=True)
