"""Shard arithmetic of the frame-sharded single-clip mode (BASELINE config 5; vista_b200/sharded.py is its only user):
contiguous frame ranges per rank and the halo neighbours of the (3,1,1) convolution."""
from __future__ import annotations

from typing import List, Optional, Tuple



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
