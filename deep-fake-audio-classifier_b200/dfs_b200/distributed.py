"""Utterance sharding across ranks (one process per GPU) and the single score all-gather.

The reference has no distributed code at all; utterances are independent through
every model, so the path shards with NO data-path collective.  The only exchange is one all-gather
of 4 bytes per utterance so that every rank can run the global min-max / EER
(src/predict_hybrid.py:81-85 and scripts/evaluation.py:7-39 act on the full score vector).

Rank r of W owns the contiguous slice [r*ceil(n/W), ...) so that concatenating the per-rank score
vectors in rank order reproduces the reference's ``shuffle=False`` utterance order
(src/predict.py:97).  Works with backend "nccl" (CUDA tensors, NVLink) and "gloo" (CPU tensors,
used by the world_size-2 CPU tests).
"""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n utterances owned by `rank`; the last ranks may be short or empty."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_sizes(n: int, world: int):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_scores(local, n_total: int | None = None):
    """All-gather per-rank 1-D score tensors (possibly ragged) into the global vector, in rank order.

    Equal shards use one ``all_gather_into_tensor``; ragged shards are padded to the largest shard and
    trimmed after the gather (still a single collective)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if n_total is None:
        sizes_t = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
        all_sizes = [torch.zeros_like(sizes_t) for _ in range(world)]
        dist.all_gather(all_sizes, sizes_t)
        sizes = [int(s.item()) for s in all_sizes]
    else:
        sizes = shard_sizes(n_total, world)
    per = max(sizes)
    buf = local
    if local.numel() != per:
        buf = torch.zeros(per, dtype=local.dtype, device=local.device)
        buf[:local.numel()] = local
    out = torch.empty(per * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf.contiguous())
    if all(s == per for s in sizes):
        return out
    return torch.cat([out[r * per:r * per + sizes[r]] for r in range(world)])
