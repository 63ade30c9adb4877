"""ctypes binding of libcsa_gpu.so (include/csa_gpu.h) -- the Python face of the C ABI.

The product path is the CUDA library built by csa_b200/csrc/Makefile.  There is NO CPU
fallback: if the library is missing, or no CUDA device is present, the calls raise.

Reference interface mirrored (fjdf/CSA, source/csamsa.c): the globals `numberofseqs`, `texts`,
`textsizes` go in; `rotations` (csamsa.h:12) and the sorted `blockslist` (csamsa.c:30) come out.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "csrc", "libcsa_gpu.so")

SET_OK, SET_NO_COMMON, SET_NO_UNIQUE, SET_DEGENERATE, SET_NONTERMINATING, SET_UNDEFINED = 0, 1, 2, 3, 4, 5
FLAG_STATS = 1
INT_MAX = 2**31 - 1


class CsaGpuError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"csa_gpu error {code}: {text}")
        self.code = code


class SetInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("nseqs", C.c_int), ("count_collected", C.c_int),
                ("count_suffixfree", C.c_int), ("count_unique", C.c_int), ("count_chains", C.c_int),
                ("nblocks", C.c_int), ("chain_is_cyclic", C.c_int), ("block_offset", C.c_longlong)]


@dataclass
class SetResult:
    """what analyzeTree() (csamsa.c:324) leaves behind for one set"""
    status: int
    rotations: Optional[np.ndarray]       # csamsa.h:12, None unless status == SET_OK
    count_collected: int
    count_suffixfree: int
    count_unique: int
    count_chains: int
    chain_is_cyclic: bool
    # the sorted blockslist (nodeslinkedlists.h:4-13)
    depth: np.ndarray = field(default=None)
    size: np.ndarray = field(default=None)
    totalsize: np.ndarray = field(default=None)
    interval: np.ndarray = field(default=None)
    next: np.ndarray = field(default=None)
    positions: np.ndarray = field(default=None)   # nblocks x nseqs
    letters: list = field(default=None)           # per block: its letters as blockLabel spells them (with_letters=True)


def _load(path):
    if not os.path.exists(path):
        raise CsaGpuError(-1, f"{path} not built (run `make -C csa_b200/csrc` or __graft_entry__.build()); "
                              "there is no CPU fallback")
    lib = C.CDLL(path)
    vp, ip, i = C.c_void_p, C.POINTER(C.c_int), C.c_int
    lib.csa_gpu_create.argtypes = [i, C.POINTER(vp)]
    lib.csa_gpu_destroy.argtypes = [vp]
    lib.csa_gpu_destroy.restype = None
    lib.csa_gpu_last_error.restype = C.c_char_p
    lib.csa_gpu_set_stream.argtypes = [vp, vp]
    lib.csa_gpu_batch_upload_flat.argtypes = [vp, i, ip, C.c_char_p, C.POINTER(C.c_longlong)]
    lib.csa_gpu_pin_host.argtypes = [vp, C.c_ulonglong]
    lib.csa_gpu_unpin_host.argtypes = [vp]
    lib.csa_gpu_batch_run.argtypes = [vp, i, C.c_uint]
    lib.csa_gpu_batch_download.argtypes = [vp, ip, C.POINTER(SetInfo)]
    lib.csa_gpu_batch_num_blocks.argtypes = [vp]
    lib.csa_gpu_batch_num_blocks.restype = C.c_longlong
    lib.csa_gpu_batch_num_positions.argtypes = [vp]
    lib.csa_gpu_batch_num_positions.restype = C.c_longlong
    lib.csa_gpu_batch_blocks.argtypes = [vp, ip, ip, ip, ip, ip, ip]
    lib.csa_gpu_batch_block_letters.argtypes = [vp, C.c_char_p, C.POINTER(C.c_longlong)]
    lib.csa_gpu_batch_block_letters.restype = C.c_longlong
    lib.csa_gpu_batch_num_suffixes.argtypes = [vp]
    lib.csa_gpu_batch_num_suffixes.restype = C.c_longlong
    lib.csa_gpu_batch_suffix_array.argtypes = [vp, C.POINTER(C.c_uint), ip]
    lib.csa_gpu_batch_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
    lib.csa_gpu_debug_rounds.argtypes = [vp, i, ip]
    lib.csa_gpu_multi_create.argtypes = [i, ip, C.POINTER(vp)]
    lib.csa_gpu_multi_destroy.argtypes = [vp]
    lib.csa_gpu_multi_destroy.restype = None
    lib.csa_gpu_multi_size.argtypes = [vp]
    lib.csa_gpu_multi_ctx.argtypes = [vp, i]
    lib.csa_gpu_multi_ctx.restype = vp
    lib.csa_gpu_multi_batch_rotations.argtypes = [vp, i, ip, C.POINTER(C.c_char_p), ip, i, C.c_uint, ip, C.POINTER(SetInfo)]
    lib.csa_gpu_shard_begin.argtypes = [vp, i, i]
    lib.csa_gpu_shard_advice.argtypes = [vp, i]
    lib.csa_gpu_shard_view.argtypes = [vp, C.POINTER(ShardInfo)]
    lib.csa_gpu_shard_finish.argtypes = [vp, i, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    lib.csa_gpu_shard_blocks_begin.argtypes = [vp, C.c_uint, C.c_void_p]
    lib.csa_gpu_shard_blocks_buffers.argtypes = [vp, C.c_uint, C.c_uint, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.csa_gpu_shard_blocks_finish.argtypes = [vp, i, C.c_uint, C.c_uint, C.c_uint]
    lib.csa_gpu_profile_enable.argtypes = [vp, i]
    lib.csa_gpu_profile_count.argtypes = [vp]
    lib.csa_gpu_profile_get.argtypes = [vp, i, C.c_char_p, i, C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double)]
    return lib


class ShardInfo(C.Structure):
    """csa_gpu_shard_info (include/csa_gpu.h)"""
    _fields_ = [("sa", C.c_void_p), ("head", C.c_void_p), ("lcp", C.c_void_p), ("left", C.c_void_p),
                ("n", C.c_ulonglong), ("bounds", C.POINTER(C.c_uint)),
                ("nleft", C.c_uint), ("left_suffixes", C.c_uint), ("min_depth", C.c_uint), ("max_group", C.c_uint),
                ("own_sort", C.c_uint)]


class ShardBlocks(C.Structure):
    """csa_gpu_shard_blocks (include/csa_gpu.h)"""
    _fields_ = [("blkrec", C.c_void_p), ("nblk", C.c_uint), ("m", C.c_uint), ("sa0", C.c_void_p), ("saidx0", C.c_void_p),
                ("lcp0", C.c_void_p), ("n0", C.c_uint), ("head_min", C.c_uint), ("tail_min", C.c_uint), ("rare", C.c_uint)]


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Batch:
    """A batch of independent sequence sets laid out flat, ready for csa_gpu_batch_upload_flat."""

    def __init__(self, sets: Sequence[Sequence[bytes]]):
        self.nsets = len(sets)
        self.set_start = np.zeros(self.nsets + 1, dtype=np.int32)
        lens = []
        for s, seqs in enumerate(sets):
            self.set_start[s + 1] = self.set_start[s] + len(seqs)
            lens.extend(len(x) for x in seqs)
        self.nseqs = int(self.set_start[-1])
        self.text_start = np.zeros(self.nseqs + 1, dtype=np.int64)
        np.cumsum(np.asarray(lens, dtype=np.int64), out=self.text_start[1:])
        self.text = b"".join(b"".join(seqs) for seqs in sets)
        self.nbases = len(self.text)

    @classmethod
    def from_arrays(cls, text: bytes, text_start: np.ndarray, set_start: np.ndarray):
        self = cls.__new__(cls)
        self.text = text
        self.text_start = np.ascontiguousarray(text_start, dtype=np.int64)
        self.set_start = np.ascontiguousarray(set_start, dtype=np.int32)
        self.nsets = len(self.set_start) - 1
        self.nseqs = int(self.set_start[-1])
        self.nbases = int(self.text_start[-1])
        return self


class RotationFinder:
    """One context = one GPU + its buffers.  find_rotations() replaces
    buildGeneralizedTree()+analyzeTree() (csamsa.c:599,610)."""

    def __init__(self, device: int = 0, lib_path: str = DEFAULT_LIB):
        self.lib = _load(lib_path)
        self.ctx = C.c_void_p()
        self._check(self.lib.csa_gpu_create(device, C.byref(self.ctx)))

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.csa_gpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle: int):
        self._check(self.lib.csa_gpu_set_stream(self.ctx, C.c_void_p(cuda_stream_handle)))

    def _check(self, rc):
        if rc != 0:
            raise CsaGpuError(rc, self.lib.csa_gpu_last_error().decode(errors="replace"))

    def pin(self, batch: "Batch"):
        """page-lock the batch's text so that upload() needs no staging copy"""
        addr = C.cast(C.c_char_p(batch.text), C.c_void_p)
        self._check(self.lib.csa_gpu_pin_host(addr, len(batch.text)))
        return addr

    def unpin(self, addr):
        self._check(self.lib.csa_gpu_unpin_host(addr))

    # ---- the three steps, separately (bench.py times run() alone and all three together) ----
    def upload(self, batch: Batch):
        self._batch = batch
        self._check(self.lib.csa_gpu_batch_upload_flat(
            self.ctx, batch.nsets, _ip(batch.set_start), batch.text,
            batch.text_start.ctypes.data_as(C.POINTER(C.c_longlong))))

    def run(self, max_interval: int = INT_MAX, flags: int = 0):
        self._check(self.lib.csa_gpu_batch_run(self.ctx, max_interval, flags))

    # ---- one batch, the suffix-array stage sharded over the ranks of a job (csa_b200/shard.py drives it) ----
    def shard_begin(self, rank: int, nranks: int):
        self._check(self.lib.csa_gpu_shard_begin(self.ctx, rank, nranks))

    def shard_advice(self, nranks: int) -> bool:
        """does sharding the uploaded batch by buckets over nranks ranks pay (include/csa_gpu.h)?"""
        return bool(self.lib.csa_gpu_shard_advice(self.ctx, nranks))

    def shard_view(self) -> ShardInfo:
        v = ShardInfo()
        self._check(self.lib.csa_gpu_shard_view(self.ctx, C.byref(v)))
        return v

    def shard_finish(self, max_interval: int, flags: int, nleft: int, left_suffixes: int, min_depth: int, max_group: int):
        self._check(self.lib.csa_gpu_shard_finish(self.ctx, max_interval, flags, nleft, left_suffixes, min_depth, max_group))

    def shard_blocks_begin(self, flags: int = 0) -> "ShardBlocks":
        b = ShardBlocks()
        self._check(self.lib.csa_gpu_shard_blocks_begin(self.ctx, flags, C.byref(b)))
        return b

    def shard_blocks_buffers(self, total_blocks: int, total_n0: int):
        p = [C.c_void_p() for _ in range(4)]
        self._check(self.lib.csa_gpu_shard_blocks_buffers(self.ctx, total_blocks, total_n0, *[C.byref(x) for x in p]))
        return [x.value for x in p]

    def shard_blocks_finish(self, max_interval: int, flags: int, total_blocks: int, total_n0: int):
        self._check(self.lib.csa_gpu_shard_blocks_finish(self.ctx, max_interval, flags, total_blocks, total_n0))

    def download(self):
        b = self._batch
        rot = np.zeros(b.nseqs, dtype=np.int32)
        info = (SetInfo * b.nsets)()
        self._check(self.lib.csa_gpu_batch_download(self.ctx, _ip(rot), info))
        return rot, info

    def blocks(self):
        nb = self.lib.csa_gpu_batch_num_blocks(self.ctx)
        ne = self.lib.csa_gpu_batch_num_positions(self.ctx)
        arrs = [np.zeros(max(nb, 1), dtype=np.int32) for _ in range(5)]
        pos = np.zeros(max(ne, 1), dtype=np.int32)
        self._check(self.lib.csa_gpu_batch_blocks(self.ctx, *[_ip(a) for a in arrs], _ip(pos)))
        return [a[:nb] for a in arrs], pos[:ne]

    def block_letters(self):
        """per block of the batch (final list order): its letters as nodeslinkedlists.c:128 blockLabel spells them"""
        nb = self.lib.csa_gpu_batch_num_blocks(self.ctx)
        off = (C.c_longlong * (nb + 1))()
        total = self.lib.csa_gpu_batch_block_letters(self.ctx, None, off)
        if total < 0:
            self._check(int(total))
        buf = C.create_string_buffer(max(int(total), 1))
        if total:
            rc = self.lib.csa_gpu_batch_block_letters(self.ctx, buf, None)
            if rc < 0:
                self._check(int(rc))
        raw = buf.raw
        return [raw[off[b]:off[b + 1]] for b in range(nb)]

    def suffix_array(self):
        n = self.lib.csa_gpu_batch_num_suffixes(self.ctx)
        sa = np.zeros(n, dtype=np.uint32)
        lcp = np.zeros(n, dtype=np.int32)
        self._check(self.lib.csa_gpu_batch_suffix_array(self.ctx, sa.ctypes.data_as(C.POINTER(C.c_uint)), _ip(lcp)))
        return sa, lcp

    def timings(self):
        ms = (C.c_float * 6)()
        launches = C.c_longlong()
        self._check(self.lib.csa_gpu_batch_timings(self.ctx, ms, C.byref(launches)))
        return list(ms), launches.value

    def debug_rounds(self, force_global: int = -1):
        r = (C.c_int * 2)()
        self._check(self.lib.csa_gpu_debug_rounds(self.ctx, force_global, r))
        return r[0], r[1]

    def profile_enable(self, on: bool):
        self._check(self.lib.csa_gpu_profile_enable(self.ctx, int(on)))

    def profile(self):
        """rows of (kernel, launches, ms, algorithmic bytes) of the last profiled run"""
        rows = []
        for j in range(self.lib.csa_gpu_profile_count(self.ctx)):
            name = C.create_string_buffer(64)
            n, ms, by = C.c_longlong(), C.c_double(), C.c_double()
            self._check(self.lib.csa_gpu_profile_get(self.ctx, j, name, 64, C.byref(n), C.byref(ms), C.byref(by)))
            rows.append((name.value.decode(), n.value, ms.value, by.value))
        return rows

    # ---- whole calls ----
    def find_rotations_batch(self, sets: Sequence[Sequence[bytes]], max_interval: int = INT_MAX,
                             flags: int = 0, with_blocks: bool = True, with_letters: bool = False) -> List[SetResult]:
        batch = sets if isinstance(sets, Batch) else Batch(sets)
        self.upload(batch)
        self.run(max_interval, flags)
        rot, info = self.download()
        out = []
        if with_blocks:
            (depth, size, total, interval, nxt), pos = self.blocks()
        letters = self.block_letters() if with_blocks and with_letters else None
        p = 0
        for s in range(batch.nsets):
            q0, q1 = int(batch.set_start[s]), int(batch.set_start[s + 1])
            inf = info[s]
            r = SetResult(inf.status, rot[q0:q1].copy() if inf.status == SET_OK else None,
                          inf.count_collected, inf.count_suffixfree, inf.count_unique, inf.count_chains,
                          bool(inf.chain_is_cyclic))
            if with_blocks:
                b0, nb, m = inf.block_offset, inf.nblocks, q1 - q0
                r.depth, r.size, r.totalsize = depth[b0:b0 + nb], size[b0:b0 + nb], total[b0:b0 + nb]
                r.interval, r.next = interval[b0:b0 + nb], nxt[b0:b0 + nb]
                r.positions = pos[p:p + nb * m].reshape(nb, m)
                if letters is not None:
                    r.letters = letters[b0:b0 + nb]
                p += nb * m
            out.append(r)
        return out

    def find_rotations(self, seqs: Sequence[bytes], max_interval: int = INT_MAX, flags: int = 0,
                       with_letters: bool = False, with_blocks: bool = True) -> SetResult:
        return self.find_rotations_batch([seqs], max_interval, flags, with_blocks=with_blocks, with_letters=with_letters)[0]


class MultiRotationFinder:
    """csa_gpu_multi_*: ONE process driving several GPUs on one batch (the suffix-array stage sharded by buckets,
    exchanged by peer copies).  Results are read from context 0, wrapped as `self.first` (a RotationFinder that
    does not own its context)."""

    def __init__(self, ngpus: int, devices: Sequence[int] = None, lib_path: str = DEFAULT_LIB):
        self.lib = _load(lib_path)
        self.m = C.c_void_p()
        dev = None if devices is None else (C.c_int * ngpus)(*devices)
        rc = self.lib.csa_gpu_multi_create(ngpus, dev, C.byref(self.m))
        if rc != 0:
            raise CsaGpuError(rc, self.lib.csa_gpu_last_error().decode(errors="replace"))
        self.first = RotationFinder.__new__(RotationFinder)
        self.first.lib = self.lib
        self.first.ctx = C.c_void_p(self.lib.csa_gpu_multi_ctx(self.m, 0))
        self.first.close = lambda: None  # owned by the multi object

    def close(self):
        if getattr(self, "m", None) and self.m.value:
            self.lib.csa_gpu_multi_destroy(self.m)
            self.m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def find_rotations_batch(self, sets: Sequence[Sequence[bytes]], max_interval: int = INT_MAX, flags: int = 0):
        """returns (rotations per sequence, SetInfo per set); blocks through self.first.blocks()"""
        batch = Batch(sets)
        flat = [x for seqs in sets for x in seqs]
        texts = (C.c_char_p * len(flat))(*flat)
        sizes = (C.c_int * len(flat))(*[len(x) for x in flat])
        rot = np.zeros(batch.nseqs, dtype=np.int32)
        info = (SetInfo * batch.nsets)()
        rc = self.lib.csa_gpu_multi_batch_rotations(self.m, batch.nsets, _ip(batch.set_start), texts, sizes, max_interval,
                                                    flags, _ip(rot), info)
        if rc != 0:
            raise CsaGpuError(rc, self.lib.csa_gpu_last_error().decode(errors="replace"))
        self.first._batch = batch
        return rot, info

