"""N>1 host logic on the CPU: two gloo ranks shard a list of sets, each runs its slice (through the
kernel single-stepper tests/emu -- logic only), rank 0 gathers; the result equals the oracle's."""
import os
import subprocess
import sys

from common import ROOT
from csa_b200.shard import shard_bounds

WORKER = r'''
import os, sys, random, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
from common import EMU_LIB, build_emu, gen_case, oracle_run
from csa_b200.api import RotationFinder
from csa_b200.shard import find_rotations_sharded
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = random.Random(123)
sets = [gen_case(rng, max_n=400)[1] for _ in range(11)]
rf = RotationFinder(lib_path=EMU_LIB)
out = find_rotations_sharded(rf, sets, rank, world, dist)
if rank == 0:
    assert len(out) == len(sets)
    for (status, rot), s in zip(out, sets):
        o = oracle_run(s)
        assert status == o["status"]
        if status == 0:
            assert rot == list(o["rotations"])
    print("SHARD_OK", world)
dist.destroy_process_group()
'''


BUCKET_WORKER = r'''
import os, sys, random
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np
import torch.distributed as dist
from common import EMU_LIB, KINDS, build_emu, gen_case, oracle_run, oracle_gsa, compare_with_oracle
from csa_b200.api import RotationFinder, Batch
from csa_b200.shard import run_bucket_sharded
import csa_b200.shard as shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = random.Random(77)
rf = RotationFinder(lib_path=EMU_LIB)
taken = {"blocks": 0, "exchange": 0}
for trial in range(24):
    # trials 0-3: ONE set (every rank sorts only its own bucket of key prefixes; trial 3: the whole set on every rank);
    # trials 4-7 and 20-23: batches of sets (first sort on every rank, buckets cut at group borders; 20-23 with the carried word sort)
    sets = [gen_case(rng, max_n=1500, kinds=KINDS if trial != 2 else ["contained", "periodic"])[1] for _ in range(1 if trial < 4 or 8 <= trial < 20 else rng.randint(2, 4))]
    mode = 8 if trial == 3 else (4 if trial % 2 and trial < 8 else 10 if trial % 2 or trial >= 20 else 0)   # 10: the word sort carried from column to column   # 4: the bucket sorts stop early and leave groups to the doubling rounds
    rf.debug_rounds(mode)
    batch = Batch(sets)
    rf.upload(batch)
    bounds = run_bucket_sharded(rf, rank, world, dist, cuda=False, shard_blocks=False)  # (buckets exchanged in full)
    assert bounds[0] == 0 and bounds[-1] == batch.nbases and bounds == sorted(bounds)
    sa, lcp = rf.suffix_array()
    off = 0
    for s in sets:   # every rank holds the whole suffix array afterwards
        n = sum(len(x) for x in s)
        osa, olcp = oracle_gsa(s)
        assert np.array_equal(sa[off:off + n].astype(np.int64) - off, osa.astype(np.int64)), (trial, "suffix array")
        assert np.array_equal(lcp[off + 1:off + n], olcp[1:]), (trial, "lcp")
        off += n
    rf.debug_rounds(0)
    ref = rf.find_rotations_batch(batch)          # the same batch on one rank alone
    rf.debug_rounds(mode)
    rf.upload(batch)
    run_bucket_sharded(rf, rank, world, dist, cuda=False)  # ONE set: the block stages on every rank's own range too
    taken[shard.last_path] += 1
    rot, info = rf.download()
    (depth, size, total, interval, nxt), pos = rf.blocks()
    for k, (s, r) in enumerate(zip(sets, ref)):
        o = oracle_run(s)
        assert info[k].status == o["status"] == r.status, (trial, info[k].status, o["status"], r.status)
        if o["status"] == 0:
            q0, q1 = int(batch.set_start[k]), int(batch.set_start[k + 1])
            assert list(rot[q0:q1]) == list(o["rotations"])
            b0, nb = info[k].block_offset, info[k].nblocks
            assert info[k].count_chains == o["count_chains"] and nb == o["nblocks"]
            for got, name in ((depth, "depth"), (size, "size"), (total, "totalsize"), (interval, "interval"), (nxt, "next")):
                assert np.array_equal(got[b0:b0 + nb], o[name]), (trial, name)
            if len(sets) == 1:
                assert np.array_equal(pos.reshape(nb, len(s)), o["positions"]), (trial, "positions")
                letters = rf.block_letters()
                assert [bytes(x) for x in letters] == [bytes(x) for x in o["letters"]], (trial, "letters")
rf.debug_rounds(0)
assert taken["blocks"] >= 6 and taken["exchange"] >= 5, taken
sys.stdout.write("BUCKET_OK_%d_of_%d %r\n" % (rank, world, taken)); sys.stdout.flush()
dist.destroy_process_group()
'''


def test_two_gloo_ranks_buckets_of_one_batch(tmp_path):
    """one batch, its suffix array built bucket by bucket on two ranks (csa_gpu_shard_*), buckets exchanged by
    gloo broadcasts: suffix array, LCP and rotations equal the oracle's on every rank"""
    from common import build_emu
    build_emu()
    w = tmp_path / "bucket_worker.py"
    w.write_text(BUCKET_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29537", str(w), ROOT],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, env=env, text=True)
    assert p.returncode == 0 and "BUCKET_OK_0_of_2" in p.stdout and "BUCKET_OK_1_of_2" in p.stdout, p.stdout[-3000:]


def test_shard_bounds():
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert shard_bounds(3, 8) == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    assert shard_bounds(0, 2) == [0, 0, 0]


def test_two_gloo_ranks(tmp_path):
    from common import build_emu
    build_emu()
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(w), ROOT],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=env, text=True)
    assert p.returncode == 0 and "SHARD_OK 2" in p.stdout, p.stdout[-2000:]


def test_stream_in_batches(emu_finder):
    import random
    from common import compare_with_oracle, gen_case, oracle_run
    from csa_b200.shard import batches_by_size, find_rotations_stream
    rng = random.Random(9)
    sets = [gen_case(rng, max_n=450)[1] for _ in range(23)]
    cuts = list(batches_by_size(sets, 3000))
    assert len(cuts) > 3 and sum(len(c) for _, c in cuts) == len(sets) and [s for s, _ in cuts] == sorted(s for s, _ in cuts)
    res = find_rotations_stream(emu_finder, sets, max_bases=3000, with_blocks=True)
    for i, (r, s) in enumerate(zip(res, sets)):
        compare_with_oracle(r, oracle_run(s), s, f"stream set {i}")
