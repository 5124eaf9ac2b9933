"""Host-side cost of one eval call of the module (python + torch + ctypes + launches) against its GPU time:
enqueue 60 calls behind a long-running kernel so that the host is never throttled by a full queue, time the
enqueue loop with perf_counter, then the GPU span with events.  Diagnostic (prints JSON)."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200"))
import torch  # noqa: E402

import gq_b200  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    reg = gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=65536, group=16, backend="cuda").to(dev).eval()
    gen = torch.Generator().manual_seed(0)
    z = torch.cat([torch.randn(64, 16, 32, 32, generator=gen) * 0.866, -1.386 + 0.3 * torch.randn(64, 16, 32, 32, generator=gen)], 1).to(dev)
    for _ in range(5):
        reg(z)
    torch.cuda.synchronize()
    n = 60
    out = {}
    # host enqueue time (GPU busy with earlier work, queue never full: 60 calls x ~9 launches)
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        reg(z)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    out["host_us_per_call"] = (t1 - t0) / n * 1e6
    out["gpu_us_per_call"] = e0.elapsed_time(e1) / n * 1e3
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        reg(z)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    print(json.dumps(out))
    print(s.getvalue()[:3500])


if __name__ == "__main__":
    main()
