"""Which synthetic batches make a step slow, and why: per batch (bench.py's seeds for the given rank range) the GPU time
of the quantize call and the pipeline's counters [rows, prefilter codewords, fallback, overflow, exact-order].  Diagnostic."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import gq_b200  # noqa: E402
from gq_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    reg = gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=65536, group=16, backend="cuda").to(dev).eval()
    seeds = [17 * r + i for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2) for i in range(20)]
    z0 = bench.synth_latents(64, 0).to(dev)
    for _ in range(5):
        reg(z0)
    cb = reg._codebook()
    out = []
    for sd in seeds:
        z = bench.synth_latents(64, sd).to(dev)
        stats = torch.zeros(8, dtype=torch.int32, device=dev)   # (the library counts in uint32 words)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eps = torch.zeros(64, 16, 32, 32, device=dev)
        e0.record()
        ops.quantize_latent(cb, z, ops.GQ_LAYOUT_BCHW, eps=eps, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        out.append((sd, round(e0.elapsed_time(e1) * 1e3, 1), stats.cpu().tolist()[:5]))
    slow = [o for o in out if o[1] > 400]
    print(json.dumps({"median_us": sorted(o[1] for o in out)[len(out) // 2], "slow": slow, "n": len(out),
                      "with_fallback": [o for o in out if o[2][2] > 0]}))


if __name__ == "__main__":
    main()
