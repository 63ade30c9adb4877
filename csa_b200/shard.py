"""Sharding of independent sequence sets over ranks (one process per GPU).

The path has no exchange step between sets, so there is no data-path collective: rank r takes a
contiguous slice of the sets, runs it on its own GPU, and the rotations are gathered once at the end
(torch.distributed, NCCL on the GPU box, gloo in the CPU tests)."""
from typing import List, Sequence

import numpy as np

from .api import RotationFinder, SetResult


def shard_bounds(nsets: int, world: int) -> List[int]:
    """first set of every rank (+ nsets): slices differ by at most one set"""
    base, extra = divmod(nsets, world)
    out = [0]
    for r in range(world):
        out.append(out[-1] + base + (1 if r < extra else 0))
    return out


def find_rotations_sharded(finder: RotationFinder, sets: Sequence[Sequence[bytes]], rank: int, world: int,
                           dist=None, flags: int = 0):
    """Every rank passes the same `sets`; returns the list of per-set (status, rotations) on rank 0
    (None elsewhere).  `dist` = torch.distributed (initialised) or None when world == 1."""
    b = shard_bounds(len(sets), world)
    mine = sets[b[rank]:b[rank + 1]]
    res: List[SetResult] = finder.find_rotations_batch(mine, flags=flags, with_blocks=False) if mine else []
    local = [(r.status, None if r.rotations is None else np.asarray(r.rotations).tolist()) for r in res]
    if world == 1:
        return local
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    return [x for part in gathered for x in part]


def batches_by_size(sets: Sequence[Sequence[bytes]], max_bases: int = 1 << 28):
    """cut a long list of sets into batches of at most `max_bases` letters (one batch's index space
    is 32 bit and ~56 B/base of HBM): yields (first set, list of sets)"""
    start, cur, tot = 0, [], 0
    for i, s in enumerate(sets):
        nb = sum(len(x) for x in s)
        if cur and tot + nb > max_bases:
            yield start, cur
            start, cur, tot = i, [], 0
        cur.append(s)
        tot += nb
    if cur:
        yield start, cur


def find_rotations_stream(finder: RotationFinder, sets: Sequence[Sequence[bytes]], max_bases: int = 1 << 28,
                          flags: int = 0, with_blocks: bool = False) -> List[SetResult]:
    """any number of independent sets through one GPU, batch after batch"""
    out: List[SetResult] = []
    for _, chunk in batches_by_size(sets, max_bases):
        out.extend(finder.find_rotations_batch(chunk, flags=flags, with_blocks=with_blocks))
    return out
