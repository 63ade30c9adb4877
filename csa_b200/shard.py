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


def _gather_var(dist, t, counts, width, dev):
    """all-gather of pieces of different lengths (counts[r] rows of `width`): padded to the longest, cut on arrival"""
    import torch
    mx = max(max(counts), 1)
    mine = torch.zeros(mx * width, dtype=torch.int32, device=dev)
    mine[:t.numel()].copy_(t)
    allb = torch.empty(len(counts) * mx * width, dtype=torch.int32, device=dev)
    if dev == "cuda":
        dist.all_gather_into_tensor(allb, mine)
    else:
        parts = [torch.empty_like(mine) for _ in counts]
        dist.all_gather(parts, mine)
        allb = torch.cat(parts)
    return torch.cat([allb[r * mx * width:r * mx * width + counts[r] * width] for r in range(len(counts))])


last_path = None  # of the last run_bucket_sharded: "blocks" (block stages on every rank's own range) | "exchange" | "replicas"


def _run_blocks_sharded(finder, rank, world, dist, max_interval, flags, v, bounds, n, m, cuda):
    """the block stages on every rank's own range (see include/csa_gpu.h, csa_gpu_shard_blocks_*).  False: some range holds
    a full-length match (rare.cuh's sets) -- the caller falls back on the full exchange."""
    import torch
    dev = "cuda" if cuda else "cpu"
    sa, lcp = _alias(v.sa, n, 4, cuda), _alias(v.lcp, n, 4, cuda)
    lo, hi = bounds[rank], bounds[rank + 1]
    # halos: the first m places of every range (suffixes and LCPs) and its last suffix, to its neighbours
    mine = torch.cat([sa[lo:lo + m], lcp[lo:lo + m], sa[hi - 1:hi]])
    if cuda:
        allh = torch.empty(world * (2 * m + 1), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allh, mine.contiguous())
        allh = allh.view(world, 2 * m + 1)
    else:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous())
        allh = torch.stack(parts)
    if rank + 1 < world:
        sa[hi:hi + m].copy_(allh[rank + 1, :m])
        lcp[hi:hi + m].copy_(allh[rank + 1, m:2 * m])
    if rank > 0:
        sa[lo - 1:lo].copy_(allh[rank - 1, 2 * m:])
    if cuda:
        torch.cuda.synchronize()
    b = finder.shard_blocks_begin(flags)
    info = torch.tensor([b.nblk, b.n0, b.head_min, b.tail_min, b.rare], dtype=torch.int64, device=dev)
    alli = [torch.zeros_like(info) for _ in range(world)]
    dist.all_gather(alli, info)
    alli = [x.tolist() for x in alli]
    if any(x[4] for x in alli):
        return False
    nblk, n0 = [x[0] for x in alli], [x[1] for x in alli]
    tb, t0 = sum(nblk), sum(n0)
    p_rec, p_sa0, p_saidx0, p_lcp0 = finder.shard_blocks_buffers(tb, t0)
    w = 2 + m
    rec = _gather_var(dist, _alias(b.blkrec, b.nblk * w, 4, cuda), nblk, w, dev)
    if tb:
        _alias(p_rec, tb * w, 4, cuda).copy_(rec)
    for src, dst in ((b.sa0, p_sa0), (b.saidx0, p_saidx0), (b.lcp0, p_lcp0)):
        _alias(dst, t0, 4, cuda).copy_(_gather_var(dist, _alias(src, b.n0, 4, cuda), n0, 1, dev))
    # the LCP of a range's first rotation of sequence 0 with the one before it: the smallest lcp between them
    lcp0 = _alias(p_lcp0, t0, 4, cuda)
    carry, at, seen = 0xFFFFFFFF, 0, False
    for r in range(world):
        if n0[r]:
            val = 0 if not seen else min(carry, alli[r][2])
            lcp0[at:at + 1].fill_(val if val < 2**31 else val - 2**32)
            carry, seen = alli[r][3], True
            at += n0[r]
        else:
            carry = min(carry, alli[r][2])
    if cuda:
        torch.cuda.synchronize()
    finder.shard_blocks_finish(max_interval, flags, tb, t0)
    return True


def run_bucket_sharded(finder: RotationFinder, rank: int, world: int, dist=None, max_interval: int = 2**31 - 1,
                       flags: int = 0, cuda: bool = True, shard_blocks: bool = True, respect_advice: bool = True):
    """The batch uploaded to `finder` on EVERY rank (the same batch), its suffix array built bucket by bucket:
    rank r orders the groups of bucket r (csa_gpu_shard_begin), the buckets -- suffix array, group heads, LCP --
    are broadcast from their owners over the job's process group (NCCL over NVLink on the GPU box, gloo in the
    CPU tests), the lists of what the bucket sorts left are concatenated, and every rank runs the rest of the
    path (csa_gpu_shard_finish).  Afterwards finder.download() / finder.blocks() give the same results on every
    rank as a single-GPU run.  Returns the bucket borders."""
    import torch
    global last_path
    if world > 1 and respect_advice and os.environ.get("CSA_SHARD_ALWAYS") != "1" and not finder.shard_advice(world):
        # one set of whole genomes on few ranks: every rank runs the whole set (the carried word sort on one GPU beats
        # the bucket sorts of up to CSA_GPU_SHARD_MIN_RANKS - 1 ranks; include/csa_gpu.h csa_gpu_shard_advice)
        last_path = "replicas"
        finder.run(max_interval, flags)
        n = finder._batch.nbases
        return [0] + [n] * world
    last_path = "exchange"
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
    batch = finder._batch
    m = int(batch.set_start[1] - batch.set_start[0])
    # ONE set of up to 64 sequences, every bucket finished by its rank, no counts asked for: the block stages run on every
    # rank's own range too and only the blocks and sequence 0's arrays travel (csa_gpu_shard_blocks_*)
    if (shard_blocks and int(v.own_sort) and batch.nsets == 1 and m <= 64 and not (flags & 1) and total == 0
            and min(bounds[r + 1] - bounds[r] for r in range(world)) >= m + 1
            and os.environ.get("CSA_SHARD_BLOCKS", "1") != "0"):
        if _run_blocks_sharded(finder, rank, world, dist, max_interval, flags, v, bounds, n, m, cuda):
            last_path = "blocks"
            return bounds
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
