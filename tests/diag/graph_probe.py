"""Single-image eval call: eager vs CUDA-graph replay (GaussianQuantRegularizer.graphed_eval)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200")); sys.path.insert(0, ROOT)
import torch, gq_b200

dev = torch.device("cuda:0")
reg = gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=65536, group=16, backend="cuda").to(dev).eval()
for b in (1, 4, 64):
    torch.manual_seed(b)
    zs = [torch.cat([torch.randn(b, 16, 32, 32) * 0.866, -1.386 + 0.3 * torch.randn(b, 16, 32, 32)], 1).to(dev) for _ in range(8)]
    run = reg.graphed_eval(zs[0])
    ok = True
    for z in zs[:4]:
        zq_e, info_e = reg(z)
        zq_e, idx_e = zq_e.clone(), info_e["indices"].clone()
        zq_g, info_g = run(z)
        ok &= torch.equal(zq_g, zq_e) and torch.equal(info_g["indices"], idx_e)
    # RNG parity: same generator state -> same zhat_noquant
    torch.manual_seed(123); _, i1 = reg(zs[1]); n1 = i1["zhat_noquant"].clone()
    torch.manual_seed(123); _, i2 = run(zs[1]); ok_rng = torch.equal(n1, i2["zhat_noquant"])
    _, i3 = run(zs[1]); fresh = not torch.equal(n1, i3["zhat_noquant"])
    res = {}
    for name, fn in (("eager", lambda z: reg(z)), ("graph", run)):
        for _ in range(10): fn(zs[0])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for k in range(300): fn(zs[k % 8])
        e1.record(); torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / 300 * 1e3, (time.perf_counter() - t0) / 300 * 1e6)
        # latency of one isolated call (host call -> result visible)
        lat = []
        for k in range(50):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(zs[k % 8]); torch.cuda.synchronize(); lat.append(time.perf_counter() - t0)
        res[name] += (sorted(lat)[25] * 1e6,)
    print(f"batch {b}: parity {ok}, rng parity {ok_rng}, fresh noise per replay {fresh}; "
          + "; ".join(f"{k}: {v[0]:.1f} us/call back-to-back (wall {v[1]:.1f}), isolated latency {v[2]:.1f} us" for k, v in res.items()), flush=True)
