"""Row sharding across the GPUs of one box (one process per GPU, torch.distributed).

The quantizer is embarrassingly parallel over rows: every row's argmax, gather and KL term
depends only on that row and the replicated codebook.  The reference itself runs the path as
pure replicas (eval.py:97-107 shards images with a DistributedSampler, nothing is exchanged;
the regularizer's lambda controller sees rank-local statistics, gaussian.py:89,101-115).  So the
data path needs no collective.  Two optional exchanges are provided for callers that want the
single-process result on every rank:

  all_gather_indices   NCCL all_gather of the int64 index shards (<= 1 MB per rank at 1 M rows)
  global_kl_stats      SUM / MIN / MAX all_reduce of the four KL partials, so that kl_loss and
                       bits-* equal the unsharded batch (a deliberate improvement over the
                       reference's rank-local lambda controller; rank-local stays the default)

Everything here works on whatever backend the process group uses (nccl on GPUs, gloo in the
CPU tests); nothing in this file touches the CUDA library.
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
    rank would have produced alone.  Shards may be ragged (global_batch % world != 0)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local.shape[0] != hi - lo:
        raise RuntimeError(f"rank {rank}: shard has {local.shape[0]} items, expected {hi - lo}")
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[: h - l] for b, (l, h) in zip(bufs, sizes)], 0)


def global_kl_stats(stats: torch.Tensor, local_groups: int, global_batch: int, lam: float, group=None) -> torch.Tensor:
    """Turn rank-local gq_kl_fwd stats into the unsharded batch's.

    stats (float32[8]) = [kl_loss, bits_mean, bits_min, bits_max, sum_kl2, sum_w_kl2, 0, 0];
    returns a new float32[4] = [kl_loss, bits_mean, bits_min, bits_max] over all ranks:
    kl_loss = lam * sum_all(w*kl2) / global_batch (gaussian.py:96-98)."""
    sums = torch.stack([stats[4].double(), stats[5].double(), torch.tensor(float(local_groups), dtype=torch.float64, device=stats.device)])
    mn = stats[2].clone()
    mx = stats[3].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out = torch.empty(4, dtype=torch.float32, device=stats.device)
    out[0] = (lam * sums[1] / float(global_batch)).float()
    out[1] = (sums[0] / sums[2]).float()
    out[2] = mn
    out[3] = mx
    return out


def global_usage_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum rank-local codebook-usage histograms (int64 [n_samples], gq_index_histogram) over all ranks,
    in place: the one exchange a sharded evaluation needs before cal_ent (eval.py:137-141).  Integer
    SUM, so the result is independent of the reduction order."""
    if hist.dtype != torch.int64:
        raise RuntimeError("hist must be int64")
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
