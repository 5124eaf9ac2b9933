"""2-rank NCCL test of the row-sharded path on GPUs (skipped on boxes with fewer than two):
every rank runs the CUDA quantizer / KL kernel on its batch shard, the NCCL exchanges of
gq_b200.distributed must reproduce what one process computes on the whole batch
(the reference's pattern: per-rank result -> all_gather, eval.py:167-201)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu
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
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import gq_b200
    from gq_b200 import distributed as gd
    from gq_b200.ops import GQ_LAYOUT_BLC
    n_samples, l, C, g = 4096, 64, 16, 16
    gen = torch.Generator().manual_seed(0)
    z = torch.cat([torch.randn(batch, l, C, generator=gen) * 0.8, -1.4 + 0.3 * torch.randn(batch, l, C, generator=gen)], 2).to(dev)
    reg = gq_b200.GaussianQuantRegularizer(format="blc", n_samples=n_samples, group=g, backend="cuda").to(dev).eval()
    zs = gd.shard_batch(z)
    _, info = reg(zs)
    full = gd.all_gather_indices(info["indices"], batch)                       # int64 over NCCL
    full_p = gd.all_gather_indices(reg.pack_indices(info["indices"]), batch)   # packed uint16 over NCCL
    hist = gd.global_usage_histogram(reg.accumulate_usage(info["indices"]))
    # train branch: rank-local KL partials -> one collective -> the unsharded batch's statistics
    eps = torch.zeros(zs.shape[0], l, C, device=dev)
    _, _, st = gq_b200.kl_penalty(zs, eps, g, GQ_LAYOUT_BLC, target_bits=12.0, lam=1.3, lam_min=0.7, lam_max=2.0)
    glob = gd.global_kl_stats(st, zs.shape[0] * l, batch, 1.3)
    # the same gather as one kernel of ours over NVLink peer memory (even splits), both wire widths, both phases of the
    # double buffer, peer stores and -- where the switch offers it -- multicast stores
    peer = {}
    if batch % world == 0:
        modes = [False]
        probe = gd.PeerIndexExchange(batch, l, n_samples, packed=True, use_multicast=False)
        if bool(getattr(probe.hdl, "has_multicast_support", False)) and int(probe.hdl.multicast_ptr) != 0:
            modes.append(True)
        for mc in modes:
            for packed in (True, False):
                for fused in (True, False):     # hand-shake inside the kernel / symmetric memory's barrier kernel behind it
                    px = probe if (not mc and packed and fused) else gd.PeerIndexExchange(
                        batch, l, n_samples, packed=packed, use_multicast=mc, fused_barrier=fused)
                    want = full_p if packed else full
                    ok = True
                    for _ in range(5):                                          # phases 0, 1, 0, 1, 0
                        got = px.gather(info["indices"])
                        ok = ok and got.dtype == want.dtype and bool(torch.equal(got.view(torch.uint8), want.view(torch.uint8)))
                    ok = ok and not px.timed_out()
                    peer[f"{px.mode}_{'packed' if packed else 'int64'}{'' if fused else '_two_kernels'}"] = ok
        pk = gd.PeerKlStatExchange()
        for _ in range(3):
            got = pk.fold(st, zs.shape[0] * l, batch, 1.3)
            peer["kl_fold"] = peer.get("kl_fold", True) and bool(torch.equal(got, glob))   # same kernel, same order: same bits
        flags = torch.tensor([float(all(peer.values()))], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)                            # every rank must have seen the right tensor
        peer["all_ranks"] = bool(flags.item() == 1.0)
    if rank == 0:
        import json
        with open(os.path.join(out_dir, "peer.json"), "w") as f:
            json.dump(peer, f)
        _, info_all = reg(z)                                                    # one process, whole batch
        _, kl_all, st_all = gq_b200.kl_penalty(z, torch.zeros(batch, l, C, device=dev), g, GQ_LAYOUT_BLC, target_bits=12.0,
                                               lam=1.3, lam_min=0.7, lam_max=2.0)
        np.savez(os.path.join(out_dir, "r0.npz"), full=full.cpu().numpy(), want=info_all["indices"].cpu().numpy(),
                 full_p=reg.unpack_indices(full_p).cpu().numpy(), packed_dtype=str(full_p.dtype),
                 hist=hist.cpu().numpy(), hist_want=reg.accumulate_usage(info_all["indices"]).cpu().numpy(),
                 glob=glob.cpu().numpy(), ref=st_all[:4].cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5])     # even and ragged split
def test_nccl_sharded_rows_match_single_process(tmp_path, batch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "r0.npz")
    assert np.array_equal(r["full"], r["want"])              # all-gathered shard indices == single-process indices
    assert np.array_equal(r["full_p"], r["want"]) and str(r["packed_dtype"]) == "torch.uint16"
    assert np.array_equal(r["hist"], r["hist_want"])
    np.testing.assert_allclose(r["glob"], r["ref"], rtol=1e-5)
    import json
    peer = json.loads((tmp_path / "peer.json").read_text())
    if batch % world == 0:
        assert peer and all(peer.values()), peer          # PeerIndexExchange == all_gather_indices, bit for bit
        assert "p2p_packed" in peer and "p2p_int64" in peer
