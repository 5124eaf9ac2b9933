"""Per-step GPU time of the headline step on every rank of a multi-GPU run (torchrun): is a rank whose 20-step
average is high uniformly slower, or does it pay for one hiccup?  Diagnostic; prints one JSON line per rank."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gq_b200  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    reg = gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=65536, group=16, backend="cuda").to(dev).eval()
    gen = torch.Generator().manual_seed(rank)
    pool = [torch.cat([torch.randn(64, 16, 32, 32, generator=gen) * 0.866, -1.386 + 0.3 * torch.randn(64, 16, 32, 32, generator=gen)], 1).to(dev)
            for _ in range(20)]
    for k in range(5):
        reg(pool[k])
    res = {}
    for trial in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n = 20
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for k in range(n):
            reg(pool[k % 20])
            ev[k + 1].record()
        torch.cuda.synchronize()
        res[f"trial{trial}"] = [round(ev[k].elapsed_time(ev[k + 1]) * 1e3, 1) for k in range(n)]
    print(json.dumps({"rank": rank, **res}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
