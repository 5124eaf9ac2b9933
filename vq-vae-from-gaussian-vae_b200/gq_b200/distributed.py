"""Row sharding across the GPUs of one box (one process per GPU, torch.distributed).

The quantizer is embarrassingly parallel over rows: every row's argmax, gather and KL term
depends only on that row and the replicated codebook.  The reference itself runs the path as
pure replicas (eval.py:97-107 shards images with a DistributedSampler, nothing is exchanged;
the regularizer's lambda controller sees rank-local statistics, gaussian.py:89,101-115).  So the
data path needs no collective.  Two optional exchanges are provided for callers that want the
single-process result on every rank:

  all_gather_indices   one NCCL all_gather_into_tensor of the index shards, int64 or the packed
                       uint16 / int32 stream (<= 1 MB resp. 256 KB per rank at 1 M rows)
  global_kl_stats      ONE all_gather of the eight gq_kl_fwd statistics per rank (32 bytes) + ONE fold kernel
                       (gq_kl_fold), so that kl_loss and bits-* equal the unsharded batch (a
                       deliberate improvement over the reference's rank-local lambda controller;
                       rank-local stays the default)

  PeerIndexExchange    the same gather as ONE kernel of ours over NVLink peer memory (gq_peer_gather_indices_sync:
                       pack + store into every rank's symmetric buffer, one multimem.st per 16 bytes where the
                       NVSwitch offers multicast, cross-rank hand-shake by the grid's last block) -- no collective
                       library on the data path; GPUs only
  PeerKlStatExchange   global_kl_stats as ONE kernel (gq_peer_kl_fold: peer stores of the 32 bytes, hand-shake, fold)

Everything else here works on whatever backend the process group uses (nccl on GPUs, gloo in the
CPU tests); only global_kl_stats on CUDA tensors calls into the CUDA library (its one-kernel fold).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n units over `world` ranks; the first n % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's contiguous slice of the batch dimension (dim 0)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def all_gather_indices(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """Concatenate per-rank index shards along dim 0 into the [global_batch, ...] tensor every
    rank would have produced alone (the pattern of eval.py:167-201: per-rank result -> all_gather).
    Works on int64 indices and on the packed uint16 / int32 stream of `pack_indices` (a quarter of
    the bytes on the wire).  Even splits are ONE `all_gather_into_tensor` straight into the result;
    ragged splits (global_batch % world != 0) pad to the largest shard."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local.shape[0] != hi - lo:
        raise RuntimeError(f"rank {rank}: shard has {local.shape[0]} items, expected {hi - lo}")
    local = local.contiguous()
    # neither NCCL nor gloo has a 16-bit integer type: ship the packed stream as its bytes
    wire = local.view(torch.uint8) if local.dtype == torch.uint16 else local
    if global_batch % world == 0:
        out = torch.empty((global_batch,) + tuple(wire.shape[1:]), dtype=wire.dtype, device=wire.device)
        dist.all_gather_into_tensor(out, wire, group=group)
    else:
        pad = wire
        if wire.shape[0] < mx:
            pad = torch.cat([wire, wire.new_zeros((mx - wire.shape[0],) + tuple(wire.shape[1:]))], 0)
        buf = torch.empty((world * mx,) + tuple(wire.shape[1:]), dtype=wire.dtype, device=wire.device)
        dist.all_gather_into_tensor(buf, pad.contiguous(), group=group)
        out = torch.cat([buf[r * mx: r * mx + (h - l)] for r, (l, h) in enumerate(sizes)], 0)
    return out.view(torch.uint16) if local.dtype == torch.uint16 else out


def global_kl_stats(stats: torch.Tensor, local_groups: int, global_batch: int, lam: float, group=None,
                    total_groups: Optional[int] = None) -> torch.Tensor:
    """Turn rank-local gq_kl_fwd stats into the unsharded batch's: ONE collective + ONE kernel.

    stats (float32[8]) = [kl_loss, bits_mean, bits_min, bits_max, sum_kl2, sum_w_kl2, 0, 0];
    returns a new float32[4] = [kl_loss, bits_mean, bits_min, bits_max] over all ranks:
    kl_loss = lam * sum_all(w*kl2) / global_batch (gaussian.py:96-98).  Every rank contributes its eight floats
    to a single all_gather (32 bytes per rank: latency-bound) and folds the world-size rows with gq_kl_fold
    (sums in double, rank order: every rank gets the same bits).  `total_groups` is the number of KL groups of
    the whole batch (global_batch * tokens * sub-codebooks); when None it is taken as local_groups scaled by
    global_batch over this rank's share, which is exact because every image contributes the same number."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if total_groups is None:
        lo, hi = shard_bounds(global_batch, rank, world)
        total_groups = local_groups // max(hi - lo, 1) * global_batch
    allv = torch.empty(world * 8, dtype=torch.float32, device=stats.device)
    dist.all_gather_into_tensor(allv, stats.contiguous(), group=group)
    out = torch.empty(4, dtype=torch.float32, device=stats.device)
    if stats.is_cuda:
        from . import _lib
        with torch.cuda.device(stats.device):
            _lib.check(_lib.lib().gq_kl_fold(allv.data_ptr(), world, float(lam), float(global_batch), float(total_groups),
                                             out.data_ptr(), torch.cuda.current_stream().cuda_stream), "gq_kl_fold")
        return out
    # host tensors (the gloo tests of this module's logic): the same fold in torch
    a = allv.view(world, 8).double()
    out[0] = (lam * a[:, 5].sum() / float(global_batch)).float()
    out[1] = (a[:, 4].sum() / float(total_groups)).float()
    out[2] = a[:, 2].min().float()
    out[3] = a[:, 3].max().float()
    return out


def global_usage_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum rank-local codebook-usage histograms (int64 [n_samples], gq_index_histogram) over all ranks,
    in place: the one exchange a sharded evaluation needs before cal_ent (eval.py:137-141).  Integer
    SUM, so the result is independent of the reduction order."""
    if hist.dtype != torch.int64:
        raise RuntimeError("hist must be int64")
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


_FLAG_BYTES = 512   # 64 uint32 flag slots + the grid's arrival counter, padded


def _symmetric_buffer(nbytes: int, group, device):
    """Zeroed symmetric allocation of nbytes + the flag block, mapped into every rank of the group; returns
    (tensor, handle).  The process-group barrier after the memset keeps a fast rank from signalling into a flag
    block its owner has not cleared yet."""
    import torch.distributed._symmetric_memory as symm
    with torch.cuda.device(device):
        buf = symm.empty(nbytes + _FLAG_BYTES, dtype=torch.uint8, device=device)
        hdl = symm.rendezvous(buf, group)
        buf.zero_()
        torch.cuda.synchronize(device)
    dist.barrier(group)
    return buf, hdl


class PeerIndexExchange:
    """all_gather of index shards as ONE kernel of this library -- no collective library on the data path.

    Every rank owns a symmetric buffer (torch.distributed._symmetric_memory: the same allocation mapped into all
    ranks of one NVLink / NVSwitch domain) holding the whole batch's indices twice (two phases) and a flag block.
    `gather(local)` launches gq_peer_gather_indices_sync: it converts this rank's int64 shard to the wire width,
    stores it at the shard's offset into every rank's buffer -- NVLink peer stores, or one `multimem.st` per 16
    bytes to the switch's multicast address (`use_multicast`, default: whenever the handle offers it) -- and its
    last block runs the cross-rank hand-shake (release store of the call's sequence number into every rank's flag
    block, acquire-poll of its own).  The return value is this rank's copy: the tensor `all_gather_indices` would
    have returned.  Calls alternate between the two phases, so a result stays valid until the call after next and
    nothing needs to be awaited before overwriting (a peer has passed the hand-shake of call t - 1 only after its
    stream finished reading the result of call t - 2).  Even batch splits only (the shard offsets are fixed at
    construction).  fused_barrier=False splits the exchange into the store kernel (gq_peer_gather_indices) and
    symmetric memory's own barrier kernel.

    packed=True ships gq_index_width(n_samples) bytes per index (uint16 up to 2^16 codewords) and returns that
    dtype (`Regularizer.unpack_indices` / `dequant_packed` read it); packed=False ships and returns int64."""

    def __init__(self, global_batch: int, items_per_image: int, n_samples: int, packed: bool = True, group=None,
                 device: Optional[torch.device] = None, use_multicast: Optional[bool] = None, fused_barrier: bool = True):
        from . import _lib
        from .ops import packed_dtype
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if global_batch % self.world:
            raise ValueError(f"PeerIndexExchange needs an even split: batch {global_batch} over {self.world} ranks")
        self.global_batch, self.items = int(global_batch), int(items_per_image)
        self.n_samples = int(n_samples)
        self.dtype = packed_dtype(n_samples) if packed else torch.int64
        self.width = torch.empty(0, dtype=self.dtype).element_size()
        self.elems = self.global_batch * self.items                      # indices of the whole batch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.phase_bytes = (self.elems * self.width + 255) // 256 * 256  # 16-byte aligned phases
        self.buf, self.hdl = _symmetric_buffer(2 * self.phase_bytes, self.group, self.device)
        mc = int(self.hdl.multicast_ptr) if bool(getattr(self.hdl, "has_multicast_support", False)) else 0
        if use_multicast is None:
            use_multicast = mc != 0
        if use_multicast and mc == 0:
            raise RuntimeError("this symmetric-memory handle has no multicast address")
        self.mc_ptr = mc if use_multicast else 0
        self.mode = "multicast" if self.mc_ptr else "p2p"
        self.fused_barrier = bool(fused_barrier)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.phase = 0
        self.seq = 0
        self._lib = _lib

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        if local.dtype != torch.int64 or not local.is_cuda:
            raise RuntimeError("PeerIndexExchange.gather takes this rank's int64 CUDA indices")
        b_local = self.global_batch // self.world
        if local.shape[0] != b_local or local.numel() != b_local * self.items:
            raise RuntimeError(f"rank {self.rank}: shard has shape {tuple(local.shape)}, expected {b_local} images of {self.items} indices")
        local = local.contiguous()
        n = local.numel()
        phase_elems = self.phase_bytes // self.width
        offset = self.phase * phase_elems + self.rank * n
        lib = self._lib
        self.seq += 1
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            if self.fused_barrier:
                lib.check(lib.lib().gq_peer_gather_indices_sync(local.data_ptr(), n, self.n_samples, self.width,
                                                                int(self.hdl.buffer_ptrs_dev), self.mc_ptr or None, self.world,
                                                                self.rank, offset, 2 * self.phase_bytes, self.seq,
                                                                self.err.data_ptr(), stream), "gq_peer_gather_indices_sync")
            else:
                lib.check(lib.lib().gq_peer_gather_indices(local.data_ptr(), n, self.n_samples, self.width,
                                                           int(self.hdl.buffer_ptrs_dev), self.mc_ptr or None, self.world, offset,
                                                           None, stream), "gq_peer_gather_indices")
                self.hdl.barrier(channel=self.phase)      # every rank's stores have landed everywhere
        lo = self.phase * self.phase_bytes
        out = self.buf[lo: lo + self.elems * self.width].view(self.dtype).view((self.global_batch,) + tuple(local.shape[1:]))
        self.phase ^= 1
        return out

    def timed_out(self) -> bool:
        """True if a hand-shake gave up waiting for a rank (synchronises; for diagnostics)."""
        return int(self.err.item()) == 2


class PeerKlStatExchange:
    """`global_kl_stats` as ONE kernel over NVLink peer memory (gq_peer_kl_fold): every rank stores its eight
    gq_kl_fwd statistics (32 bytes) into slot [rank] of every rank's symmetric buffer, runs the same cross-rank
    hand-shake as PeerIndexExchange, and folds the world-size rows it then holds exactly as gq_kl_fold does (sums
    in double, in rank order), so every rank gets the bits `global_kl_stats` returns.  Two phases alternate."""

    def __init__(self, group=None, device: Optional[torch.device] = None):
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.phase_bytes = (self.world * 32 + 255) // 256 * 256
        self.buf, self.hdl = _symmetric_buffer(2 * self.phase_bytes, self.group, self.device)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.phase = 0
        self.seq = 0
        self._lib = _lib

    def fold(self, stats: torch.Tensor, local_groups: int, global_batch: int, lam: float,
             total_groups: Optional[int] = None) -> torch.Tensor:
        if stats.dtype != torch.float32 or stats.numel() != 8 or not stats.is_cuda:
            raise RuntimeError("stats must be the float32[8] CUDA tensor gq_kl_fwd wrote")
        if total_groups is None:
            lo, hi = shard_bounds(global_batch, self.rank, self.world)
            total_groups = local_groups // max(hi - lo, 1) * global_batch
        stats = stats.contiguous()
        lib = self._lib
        out = torch.empty(4, dtype=torch.float32, device=stats.device)
        self.seq += 1
        with torch.cuda.device(self.device):
            lib.check(lib.lib().gq_peer_kl_fold(stats.data_ptr(), int(self.hdl.buffer_ptrs_dev), self.world, self.rank,
                                                self.phase * self.phase_bytes, 2 * self.phase_bytes, self.seq, float(lam),
                                                float(global_batch), float(total_groups), out.data_ptr(), self.err.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "gq_peer_kl_fold")
        self.phase ^= 1
        return out
