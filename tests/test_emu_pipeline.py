"""Kernel LOGIC on the CPU: the bodies of csa_b200/csrc/pipeline.cuh single-stepped by tests/emu
against the oracle and the reference's golden vectors.  (The CUDA build is checked by the -m gpu
tests; this file keeps the algorithm honest on a box without a GPU.)"""
import random

import numpy as np
import pytest

from common import compare_with_oracle, gen_case, oracle_gsa, oracle_run
from csa_b200 import host


def test_emu_golden_vectors(emu_finder, golden):
    for case in golden:
        seqs = [s.encode() for s in case["seqs"]]
        r = emu_finder.find_rotations(seqs, flags=1, with_letters=True)
        assert r.status == 0
        assert [r.count_collected, r.count_suffixfree, r.count_unique, r.count_chains] == case["counts"], case["name"]
        assert list(r.rotations) == case["rotations"], case["name"]
        assert host.blocks_csv(r, seqs) == case["blocks_csv"], case["name"]


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_emu_seeded_cases(emu_finder, seed):
    rng = random.Random(seed)
    for i in range(120):
        kind, seqs = gen_case(rng, max_n=1500)
        compare_with_oracle(emu_finder.find_rotations(seqs), oracle_run(seqs), seqs, f"seed {seed} case {i} {kind}")


def test_emu_batch_of_sets(emu_finder):
    rng = random.Random(77)
    sets = [gen_case(rng, max_n=500)[1] for _ in range(37)]
    res = emu_finder.find_rotations_batch(sets)
    for i, (r, s) in enumerate(zip(res, sets)):
        compare_with_oracle(r, oracle_run(s), s, f"set {i}")


def test_emu_suffix_array_and_lcp(emu_finder):
    rng = random.Random(5)
    for i in range(20):
        _, seqs = gen_case(rng, max_n=800)
        emu_finder.find_rotations(seqs)
        sa, lcp = emu_finder.suffix_array()
        osa, olcp = oracle_gsa(seqs)
        assert np.array_equal(sa.astype(np.int64), osa.astype(np.int64)), f"case {i}: suffix array"
        assert np.array_equal(lcp[1:], olcp[1:]), f"case {i}: lcp"


def test_emu_printed_counts(emu_finder, golden):
    sets = [[s.encode() for s in c["seqs"]] for c in golden]
    for c, r in zip(golden, emu_finder.find_rotations_batch(sets, flags=1)):
        assert [r.count_collected, r.count_suffixfree, r.count_unique, r.count_chains] == c["counts"], c["name"]


def test_emu_both_round_paths(emu_finder):
    rng = random.Random(4)
    sets = [gen_case(rng, max_n=1200)[1] for _ in range(25)]
    try:
        emu_finder.debug_rounds(1)
        res_g = emu_finder.find_rotations_batch(sets)
        assert emu_finder.debug_rounds()[0] == 0
        emu_finder.debug_rounds(3)
        res_q = emu_finder.find_rotations_batch(sets)
        emu_finder.debug_rounds(2)
        res_d = emu_finder.find_rotations_batch(sets)
        emu_finder.debug_rounds(5)
        res_t = emu_finder.find_rotations_batch(sets)
        emu_finder.debug_rounds(4)
        res_w4 = emu_finder.find_rotations_batch(sets)
        emu_finder.debug_rounds(6)
        res_w = emu_finder.find_rotations_batch(sets)
        assert emu_finder.debug_rounds()[0] > 0
        emu_finder.debug_rounds(10)   # the word sort, a column's order carried over to the next
        res_c = emu_finder.find_rotations_batch(sets)
        emu_finder.debug_rounds(12)   # ... with the group table of very large batches
        res_c2 = emu_finder.find_rotations_batch(sets)
        emu_finder.debug_rounds(0)
        res_0 = emu_finder.find_rotations_batch(sets)
    finally:
        emu_finder.debug_rounds(0)
    for i, (a, b, d, q, w, w4, r0, cw, s) in enumerate(zip(res_t, res_g, res_d, res_q, res_w, res_w4, res_0, res_c, sets)):
        o = oracle_run(s)
        compare_with_oracle(r0, o, s, f"free choice set {i}")
        compare_with_oracle(cw, o, s, f"carried word sort set {i}")
        compare_with_oracle(res_c2[i], o, s, f"carried word sort, unpacked group table, set {i}")
        compare_with_oracle(w, o, s, f"word sort set {i}")
        compare_with_oracle(w4, o, s, f"word sort stopped early, then group lists set {i}")
        compare_with_oracle(a, o, s, f"group-list path set {i}")
        compare_with_oracle(q, o, s, f"tile path (quadrupling) set {i}")
        compare_with_oracle(d, o, s, f"tile path (doubling) set {i}")
        compare_with_oracle(b, o, s, f"device-wide path set {i}")


def test_emu_sets_whose_tree_is_not_their_suffix_array(emu_finder):
    """csa_b200/csrc/rare.cuh: sequences that are powers w^c (identical rotations share ONE leaf,
    gencycsuffixtrees.c:507-517) and sets in which a whole rotation of the shortest sequence occurs in all others
    (a LEAF on the block list: removeSuffixNodes csamsa.c:80 and the walk of csamsa.c:147-183 follow the leaf's
    link to the next rotation).  Every status the reference can end in must turn up and be agreed on: answered
    (0), no unique block (2), walks off a leaf (3), endless block cycle (4), frees the item it stands on (5);
    counts, block list, letters and rotations bit-exact wherever there is an answer."""
    rng = random.Random(20261018)
    seen = {}
    for i in range(400):
        kind, s = gen_case(rng, max_n=400, kinds=["periodic", "contained", "ragged", "periodic", "contained"])
        o = oracle_run(s)
        seen[o["status"]] = seen.get(o["status"], 0) + 1
        compare_with_oracle(emu_finder.find_rotations(s, flags=1, with_letters=True), o, s, f"case {i} ({kind})")
        # without the counts the blocks come straight from the LCP array (k_blockfind2) and the marked sets build their own cover array
        compare_with_oracle(emu_finder.find_rotations(s, flags=0, with_letters=True), o, s, f"case {i} ({kind}), no counts")
    assert all(seen.get(st, 0) >= 3 for st in (0, 2, 3, 4, 5)), seen


def test_emu_shard_api_one_rank_and_errors(emu_finder):
    """csa_gpu_shard_begin/_view/_finish: a job of one rank gives what csa_gpu_batch_run gives; calls out of order fail"""
    from csa_b200.api import Batch, CsaGpuError
    from csa_b200.shard import run_bucket_sharded
    rng = random.Random(12)
    sets = [gen_case(rng, max_n=900)[1] for _ in range(3)]
    batch = Batch(sets)
    ref = emu_finder.find_rotations_batch(batch)
    emu_finder.upload(batch)
    with pytest.raises(CsaGpuError):
        emu_finder.shard_view()                      # before shard_begin
    with pytest.raises(CsaGpuError):
        emu_finder.shard_finish(2**31 - 1, 0, 0, 0, 0, 0)
    with pytest.raises(CsaGpuError):
        emu_finder.shard_begin(3, 2)                 # rank outside the job
    bounds = run_bucket_sharded(emu_finder, 0, 1, cuda=False)
    assert bounds == [0, batch.nbases]
    rot, info = emu_finder.download()
    for k, (r, s) in enumerate(zip(ref, sets)):
        o = oracle_run(s)
        assert info[k].status == o["status"] == r.status
        if r.status == 0:
            q0, q1 = int(batch.set_start[k]), int(batch.set_start[k + 1])
            assert list(rot[q0:q1]) == list(o["rotations"])
    with pytest.raises(CsaGpuError):
        emu_finder.shard_view()                      # the sharded run is over


@pytest.mark.parametrize("ngpus", [2, 3])
def test_emu_one_process_several_gpus(ngpus):
    """csa_gpu_multi_*: the C host's way to several GPUs (one thread per GPU, bucket exchange by peer copies), here
    over the CPU single-stepper: rotations and block lists equal the oracle's"""
    from common import EMU_LIB, build_emu
    from csa_b200.api import MultiRotationFinder
    build_emu()
    mf = MultiRotationFinder(ngpus, lib_path=EMU_LIB)
    rng = random.Random(50 + ngpus)
    try:
        for trial in range(5):
            sets = [gen_case(rng, max_n=1500)[1] for _ in range(1 if trial < 3 else 3)]
            mf.first.debug_rounds(4 if trial % 2 else 0)   # (set on GPU 0 only: the other GPUs keep the free choice)
            rot, info = mf.find_rotations_batch(sets, flags=1)
            (depth, size, total, interval, nxt), pos = mf.first.blocks()
            b0 = 0
            q = 0
            for k, s in enumerate(sets):
                o = oracle_run(s)
                assert info[k].status == o["status"], (trial, k)
                if o["status"] == 0:
                    assert list(rot[q:q + len(s)]) == list(o["rotations"])
                    nb = info[k].nblocks
                    assert np.array_equal(depth[b0:b0 + nb], o["depth"]) and np.array_equal(size[b0:b0 + nb], o["size"])
                    assert info[k].count_collected == o["count_collected"] and info[k].count_suffixfree == o["count_suffixfree"]
                b0 += info[k].nblocks
                q += len(s)
    finally:
        mf.first.debug_rounds(0)
        mf.close()



def test_emu_carried_word_sort_groups_of_hundreds(emu_finder):
    """sets of more than 32 near-identical sequences: the carried word sort walks groups of up to 256 suffixes (a CTA a root,
    k_cywalk_cta) and the blocks come straight from the LCP array for up to 256 sequences (k_blockfind2) -- the free choice
    and the forced carried sort give the suffix array and LCP array of the doubling rounds, and the oracle's answer"""
    import numpy as np
    from csa_b200.workloads import batch_sets, make_batch
    for m, n, nsets, seed in ((40, 1500, 2, 1), (100, 1200, 1, 2)):
        b = make_batch(nsets=nsets, m=m, n=n, snp=0.04, indel=0.004, seed=seed, population=True)
        out = {}
        try:
            for mode in (5, 0, 10):
                emu_finder.debug_rounds(mode)
                res = emu_finder.find_rotations_batch(b)
                out[mode] = (res,) + tuple(emu_finder.suffix_array())
        finally:
            emu_finder.debug_rounds(0)
        for mode in (0, 10):
            assert np.array_equal(out[5][1], out[mode][1]) and np.array_equal(out[5][2], out[mode][2]), (m, mode)
        for i, (r, seqs) in enumerate(zip(out[0][0], batch_sets(b))):
            compare_with_oracle(r, oracle_run(seqs), seqs, f"m={m} set {i}")
