"""world_size-2 gloo test of the row-sharding host logic (no GPU): each rank quantizes its batch
shard (with the CPU oracle standing in for the kernel), all_gather_indices / global_kl_stats must
reproduce the single-process result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_dir):
    for p in (str(ROOT), str(ROOT / "vq-vae-from-gaussian-vae_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from gq_b200 import distributed as gd
    from gq_b200_host import host_codebook
    cb, nlp = host_codebook(1024, 16)
    gen = torch.Generator().manual_seed(0)
    l, C, g = 16, 16, 16
    z = torch.cat([torch.randn(batch, l, C, generator=gen) * 0.8, -1.4 + 0.3 * torch.randn(batch, l, C, generator=gen)], 2)
    zs = gd.shard_batch(z)                          # this rank's images
    lo, hi = gd.shard_bounds(batch, rank, world)
    assert zs.shape[0] == hi - lo
    mu, lv = zs[..., :C].reshape(-1, g), zs[..., C:].reshape(-1, g)
    std = torch.exp(0.5 * lv)
    idx, _, _, _ = oracle.quantize(mu.numpy(), std.numpy(), std.log().numpy(), cb, nlp)
    local = torch.from_numpy(idx).reshape(hi - lo, l, 1)
    full = gd.all_gather_indices(local, batch)
    packed = torch.from_numpy(oracle.pack_indices(idx, 1024).view(np.int16)).view(torch.uint16).reshape(hi - lo, l, 1)
    full_packed = gd.all_gather_indices(packed, batch)          # the uint16 wire format travels as bytes
    assert full_packed.dtype == torch.uint16
    # KL partials
    r = oracle.kl(zs.numpy(), g, 10, 0.5, 1.3, 0.7, 2.0)
    kl2 = r["kl2"].astype(np.float64)
    w = np.where(kl2 > 10.5, 2.0, np.where(kl2 < 9.5, 0.7, 1.0))
    stats = torch.tensor([0, 0, r["bits_min"], r["bits_max"], kl2.sum(), (w * kl2).sum(), 0, 0], dtype=torch.float32)
    glob = gd.global_kl_stats(stats, kl2.size, batch, 1.3)
    # codebook-usage histogram of the local shard, summed over ranks
    hist = gd.global_usage_histogram(torch.from_numpy(oracle.index_histogram(idx, 1024)))
    if rank == 0:
        mu_a, lv_a = z[..., :C].reshape(-1, g), z[..., C:].reshape(-1, g)
        std_a = torch.exp(0.5 * lv_a)
        want, _, _, _ = oracle.quantize(mu_a.numpy(), std_a.numpy(), std_a.log().numpy(), cb, nlp)
        ra = oracle.kl(z.numpy(), g, 10, 0.5, 1.3, 0.7, 2.0)
        np.savez(os.path.join(out_dir, "r0.npz"), full=full.numpy(), want=want.reshape(batch, l, 1), glob=glob.numpy(),
                 full_packed=full_packed.view(torch.int16).numpy().view(np.uint16).astype(np.int64),
                 hist=hist.numpy(), hist_want=oracle.index_histogram(want, 1024),
                 ref=np.array([ra["kl_loss"], ra["bits_mean"], ra["bits_min"], ra["bits_max"]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 5])     # even and ragged split
def test_sharded_rows_match_single_process(tmp_path, batch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "r0.npz")
    assert np.array_equal(r["full"], r["want"])
    assert np.array_equal(r["full_packed"], r["want"])
    np.testing.assert_allclose(r["glob"], r["ref"], rtol=1e-5)
    assert np.array_equal(r["hist"], r["hist_want"]) and r["hist"].sum() == batch * 16


def test_shard_bounds_partition():
    from gq_b200.distributed import shard_bounds
    for n in (0, 1, 7, 64, 65536, 1000003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
