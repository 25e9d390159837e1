"""Pair-batch sharding across the GPUs of one box (one process per GPU).

The reference spreads work with ``nn.DataParallel`` (raft/train.py:71): the batch of frame pairs is
chunked over devices and every replica builds its own CorrBlock -- there is no cross-pair term
anywhere on the correlation path (raft/core/corr.py:34-35 folds the batch into the leading dim), so
the forward path needs NO collective.  These helpers carry the same partitioning over to the
one-process-per-GPU model (torch.distributed, NCCL on the box / gloo in CPU tests): ``torch.chunk``
semantics like DataParallel's scatter, an all-gather for callers that want the full batch back, and
the max-over-ranks reduction used for timing.
"""
from __future__ import annotations

import math


def shard_range(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous slice of pairs owned by ``rank`` (``torch.chunk(n, world)[rank]``; may be empty)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    per = math.ceil(n_pairs / world) if n_pairs else 0
    lo = min(rank * per, n_pairs)
    return range(lo, min(lo + per, n_pairs))


def shard(tensor, rank: int, world: int):
    """This rank's pairs of a batch-leading tensor (a view, no copy)."""
    r = shard_range(tensor.shape[0], rank, world)
    return tensor[r.start:r.stop]


def gather_pairs(local, n_pairs: int, group=None):
    """All-gather per-rank results back into batch order (only for callers that need it: the
    forward correlation path itself never communicates)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = math.ceil(n_pairs / world) if n_pairs else 0
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: len(shard_range(n_pairs, r, world))] for r, o in enumerate(out)], dim=0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
