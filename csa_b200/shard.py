"""Sharding of independent sequence sets over ranks (one process per GPU).

The path has no exchange step between sets, so there is no data-path collective: rank r takes a
contiguous slice of the sets, runs it on its own GPU, and the rotations are gathered once at the end
(torch.distributed, NCCL on the GPU box, gloo in the CPU tests)."""
import os
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


# ---- one set (or batch) too large to be quick on one GPU: the suffix-array stage sharded by buckets ----------
def _alias(ptr: int, count: int, itemsize: int, cuda: bool):
    """a torch tensor over `count` items of the library's own buffer at `ptr` (no copy): int32 or int64"""
    import torch
    if count == 0:
        return torch.empty(0, dtype=torch.int32 if itemsize == 4 else torch.int64, device="cuda" if cuda else "cpu")
    if cuda:
        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4" if itemsize == 4 else "<i8",
                                      "data": (ptr, False), "version": 2}
        return torch.as_tensor(h, device="cuda")
    import ctypes as C
    ctype = C.c_int32 if itemsize == 4 else C.c_int64
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,))
    return torch.from_numpy(arr)


def run_bucket_sharded(finder: RotationFinder, rank: int, world: int, dist=None, max_interval: int = 2**31 - 1,
                       flags: int = 0, cuda: bool = True):
    """The batch uploaded to `finder` on EVERY rank (the same batch), its suffix array built bucket by bucket:
    rank r orders the groups of bucket r (csa_gpu_shard_begin), the buckets -- suffix array, group heads, LCP --
    are broadcast from their owners over the job's process group (NCCL over NVLink on the GPU box, gloo in the
    CPU tests), the lists of what the bucket sorts left are concatenated, and every rank runs the rest of the
    path (csa_gpu_shard_finish).  Afterwards finder.download() / finder.blocks() give the same results on every
    rank as a single-GPU run.  Returns the bucket borders."""
    import torch
    finder.shard_begin(rank, world)
    v = finder.shard_view()
    bounds = [int(v.bounds[r]) for r in range(world + 1)]
    n = int(v.n)
    mine = [int(v.nleft), int(v.left_suffixes), int(v.min_depth), int(v.max_group)]
    if world == 1:
        finder.shard_finish(max_interval, flags, *mine)
        return bounds
    dev = "cuda" if cuda else "cpu"
    cnt = torch.tensor(mine, dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    allc = [c.tolist() for c in allc]
    nl = [c[0] for c in allc]
    total = sum(nl)
    # suffix array and LCP of every bucket to every rank; the group heads only when some bucket sort left groups
    # for the doubling rounds (which work on the heads)
    arrays = [_alias(ptr, n, 4, cuda) for ptr in (v.sa, v.lcp) + ((v.head,) if total else ())]
    width = max(bounds[r + 1] - bounds[r] for r in range(world))
    if cuda and os.environ.get("CSA_SHARD_EXCHANGE", "allgather") == "allgather" and width * world <= 2 * n + 1024:
        # one all-gather per array (buckets are even to within n/4096: padded to the widest), then every bucket
        # copied to its place -- one collective instead of `world` broadcasts one after the other
        mine_t = torch.empty(width, dtype=torch.int32, device=dev)
        allb = torch.empty(world * width, dtype=torch.int32, device=dev)
        for t in arrays:
            k = bounds[rank + 1] - bounds[rank]
            mine_t[:k].copy_(t[bounds[rank]:bounds[rank + 1]])
            dist.all_gather_into_tensor(allb, mine_t)
            for r in range(world):
                if r != rank and bounds[r + 1] > bounds[r]:
                    t[bounds[r]:bounds[r + 1]].copy_(allb[r * width:r * width + bounds[r + 1] - bounds[r]])
    else:
        for t in arrays:
            for r in range(world):
                if bounds[r + 1] > bounds[r]:
                    dist.broadcast(t[bounds[r]:bounds[r + 1]], src=r)
    if total:
        left = _alias(v.left, max(total, nl[rank]), 8, cuda)
        tmp = torch.empty(total, dtype=torch.int64, device=dev)
        off = 0
        for r in range(world):
            if nl[r]:
                seg = tmp[off:off + nl[r]]
                if r == rank:
                    seg.copy_(left[:nl[r]])
                dist.broadcast(seg, src=r)
                off += nl[r]
        left[:total].copy_(tmp)
    if cuda:
        torch.cuda.synchronize()
    finder.shard_finish(max_interval, flags, total, sum(c[1] for c in allc), min(c[2] for c in allc), max(c[3] for c in allc))
    return bounds
