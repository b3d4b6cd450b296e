"""Multi-GPU plumbing (host logic only; torch.distributed is the transport).

Round 1 ships clip-level sharding: each rank samples whole clips, no data-path collective (SURVEY §8e:
independent units -> weak scaling).  The helpers for the frame-sharded single-clip mode of BASELINE
config 5 (contiguous frame ranges, halo neighbours, gather layout) live here too so that the gloo tests
pin their arithmetic before the CUDA side lands.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_units(n_units: int, world: int, rank: int) -> range:
    """Contiguous, balanced assignment of independent units (clips) to ranks."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def frame_shards(num_frames: int, world: int) -> List[Tuple[int, int]]:
    """[start, stop) frame range of every rank: contiguous, sizes differ by at most one (25 frames / 8 ranks ->
    4,3,3,3,3,3,3,3 — SURVEY §8e)."""
    out = []
    for r in range(world):
        rg = shard_units(num_frames, world, r)
        out.append((rg.start, rg.stop))
    return out


def halo_neighbours(rank: int, world: int) -> Tuple[Optional[int], Optional[int]]:
    """Ranks owning frame t-1 of our first frame and t+1 of our last frame (None at the clip ends: zero padding of
    the (3,1,1) convolution, openaimodel.py:190-193)."""
    return (rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None)


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule of the bench contract: report the slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frames(local: torch.Tensor, num_frames: int) -> torch.Tensor:
    """All-gather of a frame-sharded tensor (frames on dim 0, uneven shards) into the full clip on every rank —
    the layout of the temporal-attention K/V gather and of the final latent gather before decode."""
    world, rank = dist.get_world_size(), dist.get_rank()
    shards = frame_shards(num_frames, world)
    assert local.shape[0] == shards[rank][1] - shards[rank][0]
    pad = max(b - a for a, b in shards)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, shards)], dim=0)


def reduce_group_stats(sums: torch.Tensor) -> torch.Tensor:
    """Sum of per-rank GroupNorm partial sums ([..., 2] = sum, sum of squares) in rank order — the temporal
    GroupNorm statistic spans all frames (video_model.py:67-72)."""
    world = dist.get_world_size()
    parts = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(parts, sums)
    total = parts[0].clone()
    for p in parts[1:]:          # fixed order: every rank computes bit-identical statistics
        total += p
    return total
