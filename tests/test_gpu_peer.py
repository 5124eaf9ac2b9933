"""The peer-memory exchange kernels on ONE GPU, through the C ABI: `world` buffers in the same device memory stand
in for the ranks' peer-mapped buffers, and the `world` "ranks" are kernels on separate streams of one process.  What
is checked is the kernels' own logic -- shard offsets, wire widths, tails, the cross-rank hand-shake (sequence
numbers, phases) and the fold -- against numpy; tests/test_gpu_distributed.py covers the real thing (two GPUs,
torch symmetric memory, NVLink peer and multicast stores) against the NCCL gather."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

FLAG_BYTES = 512


def _lib():
    from gq_b200 import _lib
    return _lib


def _ptr_array(tensors):
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device="cuda")


@pytest.mark.parametrize("width,N", [(2, 65536), (4, 262144), (8, 65536)])
@pytest.mark.parametrize("n_local", [4096, 1003])       # vector path / ragged tail + unaligned shard offsets
def test_peer_gather_store_kernel(width, N, n_local):
    """gq_peer_gather_indices: every "rank's" shard lands at its offset in every buffer, converted to the wire width."""
    L = _lib()
    world = 3
    dt = {2: np.uint16, 4: np.uint32, 8: np.int64}[width]
    bufs = [torch.zeros(world * n_local * width + 64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ptrs = _ptr_array(bufs)
    rng = np.random.default_rng(0)
    shards = [rng.integers(0, N, n_local).astype(np.int64) for _ in range(world)]
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for r in range(world):
        idx = torch.from_numpy(shards[r]).cuda()
        L.check(L.lib().gq_peer_gather_indices(idx.data_ptr(), n_local, N, width, ptrs.data_ptr(), None, world, r * n_local,
                                               flag.data_ptr(), torch.cuda.current_stream().cuda_stream), "gq_peer_gather_indices")
    torch.cuda.synchronize()
    want = np.concatenate(shards).astype(dt)
    for b in bufs:
        got = b.cpu().numpy()[: world * n_local * width].view(dt)
        assert np.array_equal(got, want)
    assert int(flag.item()) == 0
    # out-of-range indices are flagged like gq_pack_indices does
    bad = torch.tensor([N, 1, 2, 3, 4, 5, 6, 7], dtype=torch.int64, device="cuda")
    L.check(L.lib().gq_peer_gather_indices(bad.data_ptr(), 8, N, width, ptrs.data_ptr(), None, world, 0, flag.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "gq_peer_gather_indices")
    assert int(flag.item()) == 1


def test_peer_gather_with_in_kernel_handshake():
    """gq_peer_gather_indices_sync with three "ranks" as three streams: nobody leaves before everybody has stored, and
    sequence numbers / phases carry over several calls."""
    L = _lib()
    world, n_local, N, width = 3, 8192, 65536, 2
    phase_bytes = world * n_local * width
    bufs = [torch.zeros(2 * phase_bytes + FLAG_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ptrs = _ptr_array(bufs)
    streams = [torch.cuda.Stream() for _ in range(world)]
    errs = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
    rng = np.random.default_rng(1)
    torch.cuda.synchronize()
    for call in range(1, 5):
        phase = (call - 1) & 1
        shards = [rng.integers(0, N, n_local).astype(np.int64) for _ in range(world)]
        dev_shards = [torch.from_numpy(s).cuda() for s in shards]
        snaps = []
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                L.check(L.lib().gq_peer_gather_indices_sync(dev_shards[r].data_ptr(), n_local, N, width, ptrs.data_ptr(), None,
                                                            world, r, phase * (phase_bytes // width) + r * n_local,
                                                            2 * phase_bytes, call, errs[r].data_ptr(), streams[r].cuda_stream),
                        "gq_peer_gather_indices_sync")
                # stream-ordered right behind the exchange: what this "rank" would read
                snaps.append(bufs[r][phase * phase_bytes: (phase + 1) * phase_bytes].clone())
        torch.cuda.synchronize()
        want = np.concatenate(shards).astype(np.uint16)
        for r in range(world):
            assert np.array_equal(snaps[r].cpu().numpy().view(np.uint16), want), (call, r)
            assert int(errs[r].item()) == 0


def test_peer_handshake_alone_orders_streams():
    """n = 0: the entry point is a pure cross-rank barrier.  "Rank" 1 is held back by a long kernel on its stream;
    "rank" 0 must not get past its hand-shake before rank 1 has made the call."""
    L = _lib()
    world = 2
    bufs = [torch.zeros(256 + FLAG_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ptrs = _ptr_array(bufs)
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
    errs = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
    marker = torch.zeros(1, dtype=torch.int32, device="cuda")
    seen = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    big = torch.randn(4096, 4096, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        for _ in range(20):
            big = big @ big * 1e-3          # ~ms of work ahead of rank 1's call
        marker.fill_(1)
        L.check(L.lib().gq_peer_gather_indices_sync(None, 0, 65536, 2, ptrs.data_ptr(), None, world, 1, 0, 256, 1,
                                                    errs[1].data_ptr(), s1.cuda_stream), "gq_peer_gather_indices_sync")
    with torch.cuda.stream(s0):
        L.check(L.lib().gq_peer_gather_indices_sync(None, 0, 65536, 2, ptrs.data_ptr(), None, world, 0, 0, 256, 1,
                                                    errs[0].data_ptr(), s0.cuda_stream), "gq_peer_gather_indices_sync")
        seen.copy_(marker)                  # ordered behind rank 0's hand-shake
    torch.cuda.synchronize()
    assert int(seen.item()) == 1 and int(errs[0].item()) == 0 and int(errs[1].item()) == 0


def test_peer_handshake_times_out_instead_of_hanging():
    """A rank whose peers never make the call gives up after a few seconds and reports it (err_flag = 2)."""
    L = _lib()
    world = 2
    bufs = [torch.zeros(1024 + FLAG_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ptrs = _ptr_array(bufs)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    idx = torch.arange(16, dtype=torch.int64, device="cuda")
    L.check(L.lib().gq_peer_gather_indices_sync(idx.data_ptr(), 16, 65536, 2, ptrs.data_ptr(), None, world, 0, 0, 1024, 1,
                                                err.data_ptr(), torch.cuda.current_stream().cuda_stream), "gq_peer_gather_indices_sync")
    torch.cuda.synchronize()
    assert int(err.item()) == 2


def test_peer_kl_fold_equals_gq_kl_fold():
    """gq_peer_kl_fold on four "ranks" (streams): every rank ends with the bits gq_kl_fold computes from the gathered rows."""
    L = _lib()
    world = 4
    phase_bytes = 256
    bufs = [torch.zeros(2 * phase_bytes + FLAG_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ptrs = _ptr_array(bufs)
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(2)
    for call in range(1, 4):
        phase = (call - 1) & 1
        stats = rng.normal(size=(world, 8)).astype(np.float32)
        stats[:, 4:6] = np.abs(stats[:, 4:6]) * 1e4
        if call == 3:
            stats[1, 2] = np.nan                      # a NaN minimum of one shard propagates (torch.min semantics)
        dev_stats = [torch.from_numpy(stats[r]).cuda() for r in range(world)]
        outs = [torch.empty(4, dtype=torch.float32, device="cuda") for _ in range(world)]
        errs = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                L.check(L.lib().gq_peer_kl_fold(dev_stats[r].data_ptr(), ptrs.data_ptr(), world, r, phase * phase_bytes,
                                                2 * phase_bytes, call, 1.3, 64.0, 65536.0, outs[r].data_ptr(), errs[r].data_ptr(),
                                                streams[r].cuda_stream), "gq_peer_kl_fold")
        torch.cuda.synchronize()
        want = torch.empty(4, dtype=torch.float32, device="cuda")
        allv = torch.from_numpy(stats.reshape(-1)).cuda()
        L.check(L.lib().gq_kl_fold(allv.data_ptr(), world, 1.3, 64.0, 65536.0, want.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), "gq_kl_fold")
        torch.cuda.synchronize()
        for r in range(world):
            assert int(errs[r].item()) == 0
            assert np.array_equal(outs[r].cpu().numpy().view(np.uint32), want.cpu().numpy().view(np.uint32)), (call, r)


def test_peer_entry_points_reject_bad_arguments():
    L = _lib()
    buf = torch.zeros(1024, dtype=torch.uint8, device="cuda")
    ptrs = _ptr_array([buf])
    idx = torch.zeros(8, dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    lib = L.lib()
    assert lib.gq_peer_gather_indices(idx.data_ptr(), 8, 65536, 4, ptrs.data_ptr(), None, 1, 0, None, s) != 0      # width != index width
    assert lib.gq_peer_gather_indices(idx.data_ptr(), 8, 65536, 2, None, None, 1, 0, None, s) != 0                  # no buffers
    assert lib.gq_peer_gather_indices_sync(idx.data_ptr(), 8, 65536, 2, ptrs.data_ptr(), None, 1, 0, 0, 512, 0, None, s) != 0   # seq 0
    assert lib.gq_peer_gather_indices_sync(idx.data_ptr(), 8, 65536, 2, ptrs.data_ptr(), None, 1, 1, 0, 512, 1, None, s) != 0   # rank >= world
    assert lib.gq_peer_kl_fold(None, ptrs.data_ptr(), 1, 0, 0, 512, 1, 1.0, 1.0, 1.0, buf.data_ptr(), None, s) != 0
    assert b"gq_peer_kl_fold" in lib.gq_last_error()
