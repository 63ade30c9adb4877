"""-m gpu: the product (libcsa_gpu.so on cuda:0, through the C ABI) against the oracle and the
reference's golden vectors.  Bit-exact: rotations, the whole sorted block list, counts, status."""
import random

import numpy as np
import pytest

from common import compare_with_oracle, gen_case, oracle_gsa, oracle_run
from csa_b200 import host
from csa_b200.workloads import batch_sets, workload_batch

pytestmark = pytest.mark.gpu


def test_gpu_golden_vectors(gpu_finder, golden):
    """BASELINE.json configs[0] (Primates.txt) and configs[1] (Mammals.txt) + 80 synthetic sets:
    outputs of the unmodified reference binary"""
    sets = [[s.encode() for s in c["seqs"]] for c in golden]
    res = gpu_finder.find_rotations_batch(sets, flags=1, with_letters=True)
    for c, seqs, r in zip(golden, sets, res):
        assert r.status == 0, c["name"]
        assert [r.count_collected, r.count_suffixfree, r.count_unique, r.count_chains] == c["counts"], c["name"]
        assert list(r.rotations) == c["rotations"], c["name"]
        assert host.blocks_csv(r, seqs) == c["blocks_csv"], c["name"]
    # and one set at a time (the reference's own call shape)
    for c, seqs in list(zip(golden, sets))[:6]:
        r = gpu_finder.find_rotations(seqs)
        assert list(r.rotations) == c["rotations"], c["name"]


def test_gpu_where_the_reference_does_not_finish(gpu_finder, golden_edge):
    """tests/golden/golden_edge.json.gz: the reference died or never returned; the CUDA path classifies as the oracle does
    (a ring in a printed chain when the reference got as far as -Rotated.fasta, else status 3 / 4 / 5)"""
    import hashlib
    sets = [[s.encode() for s in c["seqs"]] for c in golden_edge]
    res = gpu_finder.find_rotations_batch(sets, flags=1)
    for c, seqs, r in zip(golden_edge, sets, res):
        o = oracle_run(seqs)
        compare_with_oracle(r, o, seqs, c["name"])
        if c["rotated_sha256"] is not None:
            assert r.status == 0, c["name"]
            assert hashlib.sha256(host.rotated_fasta(c["descs"], seqs, r.rotations)).hexdigest() == c["rotated_sha256"], c["name"]
            assert any(host.chain_is_ring(r, b) for b in range(len(r.depth)) if r.totalsize[b] != -1), c["name"]
        else:
            assert r.status in ((3, 4) if c["outcome"] == "hangs" else (3, 5)), (c["name"], r.status)


def test_gpu_sets_whose_tree_is_not_their_suffix_array(gpu_finder):
    """csa_b200/csrc/rare.cuh on the device (see tests/test_emu_pipeline.py for what these sets are): powers w^c,
    one sequence wholly inside all others; mixed into one batch with ordinary sets so that marked and unmarked sets
    run side by side in the same launches.  Bit-exact with the oracle, every status, counts, letters."""
    rng = random.Random(20261018)
    cases = [gen_case(rng, max_n=400, kinds=["periodic", "contained", "ragged", "periodic", "contained", "variants"]) for _ in range(600)]
    sets = [c[1] for c in cases]
    res = gpu_finder.find_rotations_batch(sets, flags=1, with_letters=True)
    seen = {}
    oras = [oracle_run(s) for s in sets]
    for i, (r, o, s) in enumerate(zip(res, oras, sets)):
        seen[o["status"]] = seen.get(o["status"], 0) + 1
        compare_with_oracle(r, o, s, f"set {i} {cases[i][0]}")
    assert all(seen.get(st, 0) >= 3 for st in (0, 2, 3, 4, 5)), seen
    # without the counts: blocks straight from the LCP array (k_blockfind2), the marked sets build their own cover array
    res0 = gpu_finder.find_rotations_batch(sets, flags=0, with_letters=True)
    for i, (r, o, s) in enumerate(zip(res0, oras, sets)):
        compare_with_oracle(r, o, s, f"set {i} {cases[i][0]} (no counts)")
    # one marked set alone (the reference's call shape), each kind
    for i in (0, 1, 2, 3, 4, 5, 6, 7):
        compare_with_oracle(gpu_finder.find_rotations(sets[i], flags=1, with_letters=True), oras[i], sets[i], f"alone {i}")


def test_gpu_first_sequence_the_longest(gpu_finder):
    """a two-sequence set whose FIRST sequence is by far the longer: the sequence-0 tree has 2 N0 > N nodes and sorts in
    the suffix array's buffers (round-1 advisor finding: they were sized for N)"""
    rng = random.Random(77)
    for n0, n1 in ((3000, 2000), (5000, 600), (40000, 9000)):
        a = bytes(rng.choice(b"ACGT") for _ in range(n0))
        b = bytearray(a[n0 // 6:n0 // 6 + n1])
        for i in range(0, len(b), 97):
            b[i] = rng.choice(b"ACGT")
        s = [a, bytes(b)]
        compare_with_oracle(gpu_finder.find_rotations(s, flags=1, with_letters=True), oracle_run(s), s, f"{n0}/{n1}")


@pytest.mark.parametrize("seed", [21, 22, 23, 24])
def test_gpu_seeded_sets_vs_oracle(gpu_finder, seed):
    rng = random.Random(seed)
    cases = [gen_case(rng, max_n=2500) for _ in range(150)]
    sets = [c[1] for c in cases]
    res = gpu_finder.find_rotations_batch(sets)
    for i, (r, s) in enumerate(zip(res, sets)):
        compare_with_oracle(r, oracle_run(s), s, f"seed {seed} set {i} {cases[i][0]}")
    res = gpu_finder.find_rotations_batch(sets[:40], flags=1, with_letters=True)
    for i, (r, s) in enumerate(zip(res, sets)):
        compare_with_oracle(r, oracle_run(s), s, f"seed {seed} set {i} {cases[i][0]} (counts and letters)")


def test_gpu_suffix_array_and_lcp(gpu_finder):
    rng = random.Random(5)
    for i in range(12):
        _, seqs = gen_case(rng, max_n=2500)
        gpu_finder.find_rotations(seqs)
        sa, lcp = gpu_finder.suffix_array()
        osa, olcp = oracle_gsa(seqs)
        assert np.array_equal(sa.astype(np.int64), osa.astype(np.int64)), f"case {i}: suffix array"
        assert np.array_equal(lcp[1:], olcp[1:]), f"case {i}: lcp"


def check_block_properties(seqs, r):
    """size-independent properties of a correct answer: every block occurs at its position in every
    sequence, the blocks of the head chain are in one circular order everywhere, and the cut points
    are the positions of the head block"""
    assert r.status == 0
    m = len(seqs)
    for b in range(len(r.depth)):
        d = int(r.depth[b])
        ref = None
        for k in range(m):
            p, s = int(r.positions[b][k]), seqs[k]
            w = (s + s)[p:p + d]
            norm = bytes(c if c in b"ACGT" else 45 for c in w)
            ref = norm if ref is None else ref
            assert norm == ref, f"block {b} differs in sequence {k}"
    assert list(r.rotations) == [int(x) for x in r.positions[0]]


@pytest.mark.parametrize("name,nsets,oracle_sets", [("mammals", 24, 24), ("sets32", 6, 2), ("variants256", 1, 0)])
def test_gpu_baseline_shapes(gpu_finder, name, nsets, oracle_sets):
    """the shapes of BASELINE.json configs[1..3] at full sequence length: properties on every set,
    the oracle on as many sets as it finishes in seconds"""
    batch = workload_batch(name, nsets, seed=99)
    sets = batch_sets(batch)
    res = gpu_finder.find_rotations_batch(batch)
    for i, (r, s) in enumerate(zip(res, sets)):
        check_block_properties(s, r)
        if i < oracle_sets:
            compare_with_oracle(r, oracle_run(s), s, f"{name} set {i}")


def test_gpu_rotation_equivariance(gpu_finder):
    """rotating the inputs moves the cut points with them (mod n) when no tie-break is involved"""
    batch = workload_batch("sets32", 1, seed=5)
    seqs = batch_sets(batch)[0]
    r0 = gpu_finder.find_rotations(seqs)
    rng = random.Random(1)
    shifts = [rng.randrange(len(s)) for s in seqs]
    moved = [s[sh:] + s[:sh] for s, sh in zip(seqs, shifts)]
    r1 = gpu_finder.find_rotations(moved)
    assert sorted(r0.depth) == sorted(r1.depth)
    head0 = [(s + s)[p:p + int(r0.depth[0])] for s, p in zip(seqs, r0.rotations)]
    head1 = [(s + s)[p:p + int(r1.depth[0])] for s, p in zip(moved, r1.rotations)]
    if head0[0] == head1[0]:
        assert [(int(p) + sh) % len(s) for p, sh, s in zip(r1.rotations, shifts, seqs)] == [int(p) for p in r0.rotations]


def test_gpu_edge_cases(gpu_finder):
    cases = [
        [b"ACGTACGTTT", b"ACGTACGAAT"],                      # tiny
        [b"AC", b"CA" + b"G"],                                  # shortest allowed
        [b"ACGTN" * 7 + b"A", b"ACGTN" * 7 + b"C"],             # IUPAC letters
        [b"A" * 50 + b"C", b"A" * 40 + b"CG"],                  # low complexity
        [b"ACGGTCA" * 9, b"TTGACCA" * 8, b"GGGTTTCA" * 5],      # periodic, ragged lengths
    ]
    res = gpu_finder.find_rotations_batch(cases)
    for i, (r, s) in enumerate(zip(res, cases)):
        compare_with_oracle(r, oracle_run(s), s, f"edge {i}")


def test_gpu_errors(gpu_finder):
    from csa_b200.api import CsaGpuError
    with pytest.raises(CsaGpuError):
        gpu_finder.find_rotations([b"ACGT"])  # one sequence: the path needs two (csamsa.c:533)
    with pytest.raises(CsaGpuError):
        gpu_finder.find_rotations([b"ACGT", b""])


def test_gpu_tile_rounds_equal_device_wide_rounds(gpu_finder):
    """the one-pass tile rounds (k_refine4: h -> 4h, k_refine: h -> 2h) and the device-wide radix round
    give the same suffix array"""
    out = {}
    try:
        for name, batch in (("sets32", workload_batch("sets32", 3, seed=3)), ("mammals", workload_batch("mammals", 4, seed=4)),
                            ("variants256", workload_batch("variants256", 1, seed=5))):
            # 0 free choice (word sort first), 5 free choice among the doubling rounds (group lists while groups are
            # small), 4 word sort stopped after two words + doubling, 3 tile rounds with quadrupling, 2 tile rounds
            # with doubling only (+ text-order LCP), 1 device-wide rounds only
            # 6 word sort whatever the groups look like (0 picks it only for small groups); 10 the same, a column's order
            # carried over to the next (what sets of whole genomes take)
            for mode in (0, 6, 5, 4, 3, 2, 1, 10, 12):
                gpu_finder.debug_rounds(mode)
                res = gpu_finder.find_rotations_batch(batch)
                sa, lcp = gpu_finder.suffix_array()
                out[(name, mode)] = (res, sa, lcp, gpu_finder.debug_rounds())
    finally:
        gpu_finder.debug_rounds(0)
    for name in ("sets32", "mammals", "variants256"):
        r0, sa0, lcp0, rounds0 = out[(name, 0)]
        assert rounds0[0] > 0
        assert out[(name, 1)][3][0] == 0 and out[(name, 1)][3][1] > 0
        assert out[(name, 2)][3][0] >= out[(name, 3)][3][0]  # doubling needs at least as many rounds as quadrupling
        for mode in (1, 2, 3, 4, 5, 6, 10, 12):
            r, sa, lcp, _ = out[(name, mode)]
            assert np.array_equal(sa0, sa) and np.array_equal(lcp0, lcp), (name, mode)
            for a, b in zip(r0, r):
                assert np.array_equal(a.rotations, b.rotations) and np.array_equal(a.positions, b.positions)


def test_gpu_large_groups_fall_back(gpu_finder):
    """low-complexity sequences make groups larger than a tile: those rounds take the device-wide path"""
    rng = random.Random(8)
    seqs = []
    for _ in range(3):
        s = bytearray(b"A" * 6000)
        for _ in range(40):
            s[rng.randrange(len(s))] = rng.choice(b"CGT")
        seqs.append(bytes(s))
    r = gpu_finder.find_rotations(seqs)
    _, glob = gpu_finder.debug_rounds()
    assert glob > 0
    compare_with_oracle(r, oracle_run(seqs), seqs, "polyA")


def test_gpu_printed_counts(gpu_finder, golden):
    """csamsa.c:332,338: "nodes found" / "nodes left" (CSA_GPU_FLAG_STATS) against the reference's stdout"""
    sets = [[s.encode() for s in c["seqs"]] for c in golden]
    res = gpu_finder.find_rotations_batch(sets, flags=1)
    for c, r in zip(golden, res):
        assert [r.count_collected, r.count_suffixfree, r.count_unique, r.count_chains] == c["counts"], c["name"]
    rng = random.Random(31)
    cases = [gen_case(rng, max_n=2000)[1] for _ in range(120)]
    for i, (r, s) in enumerate(zip(gpu_finder.find_rotations_batch(cases, flags=1), cases)):
        assert r.count_collected >= 0
        compare_with_oracle(r, oracle_run(s), s, f"stats case {i}")


def test_gpu_upload_from_page_locked_buffer(gpu_finder):
    """csa_gpu_pin_host: the upload reads the caller's page-locked buffer in place; same answers"""
    batch = workload_batch("mammals", 5, seed=17)
    r0 = gpu_finder.find_rotations_batch(batch)
    h = gpu_finder.pin(batch)
    try:
        r1 = gpu_finder.find_rotations_batch(batch)
    finally:
        gpu_finder.unpin(h)
    for a, b in zip(r0, r1):
        assert a.status == b.status and np.array_equal(a.rotations, b.rotations) and np.array_equal(a.positions, b.positions)


def test_gpu_many_tiny_sets(gpu_finder):
    """thousands of sets in one batch (per-set tables, segmented first sort, one chain CTA per set)"""
    rng = random.Random(12)
    sets = []
    for _ in range(3000):
        n = rng.randint(30, 90)
        base = [rng.choice("ACGT") for _ in range(n)]
        seqs = []
        for _ in range(rng.randint(2, 4)):
            s = [c if rng.random() > 0.03 else rng.choice("ACGT") for c in base]
            r = rng.randrange(n)
            seqs.append("".join(s[r:] + s[:r]).encode())
        from common import drop_rotation_duplicates
        seqs = drop_rotation_duplicates(seqs)
        if len(seqs) >= 2:
            sets.append(seqs)
    res = gpu_finder.find_rotations_batch(sets)
    assert len(res) == len(sets)
    for i in range(0, len(sets), 7):
        compare_with_oracle(res[i], oracle_run(sets[i]), sets[i], f"tiny set {i}")


def test_gpu_large_batch_of_config3_sets(gpu_finder):
    """BASELINE configs[3] at full set size, 160 sets (85 M bases) in one batch: every set answered, the
    block properties hold on a sample, three sets checked against the oracle"""
    batch = workload_batch("sets32", 160, seed=2026)
    res = gpu_finder.find_rotations_batch(batch)
    assert len(res) == 160 and all(r.status == 0 for r in res)
    sets = batch_sets(batch)
    for i in range(0, 160, 16):
        check_block_properties(sets[i], res[i])
    for i in (0, 77, 159):
        compare_with_oracle(res[i], oracle_run(sets[i]), sets[i], f"set {i}")


def test_gpu_bucket_path_one_rank(gpu_finder):
    """csa_gpu_shard_begin / _finish with a job of one rank (the whole suffix array is one bucket) give what
    csa_gpu_batch_run gives; the N > 1 exchange is covered by tests/test_sharding.py on gloo"""
    from csa_b200.shard import run_bucket_sharded
    for name, nsets in (("mammals", 3), ("sets32", 2)):
        batch = workload_batch(name, nsets, seed=11)
        ref = gpu_finder.find_rotations_batch(batch)
        sa0, lcp0 = gpu_finder.suffix_array()
        gpu_finder.upload(batch)
        bounds = run_bucket_sharded(gpu_finder, 0, 1)
        assert bounds == [0, batch.nbases]
        sa, lcp = gpu_finder.suffix_array()
        assert np.array_equal(sa0, sa) and np.array_equal(lcp0[1:], lcp[1:])
        rot, info = gpu_finder.download()
        for k, r in enumerate(ref):
            q0, q1 = int(batch.set_start[k]), int(batch.set_start[k + 1])
            assert info[k].status == r.status
            if r.status == 0:
                assert np.array_equal(rot[q0:q1], r.rotations)


def test_gpu_long_block_lists(gpu_finder):
    """a set with thousands of blocks (bacterial-style: few long genomes) takes k_chain_big -- walk order in shared
    memory, sums in parallel -- and must leave exactly what the literal one-thread walk (mode 7) and the oracle leave"""
    from csa_b200.workloads import make_batch, batch_sets
    batch = make_batch(2, 4, 300_000, 0.01, 0.001, seed=21, population=True)
    try:
        gpu_finder.debug_rounds(7)
        lit = gpu_finder.find_rotations_batch(batch)
    finally:
        gpu_finder.debug_rounds(0)
    res = gpu_finder.find_rotations_batch(batch)
    for a, b in zip(lit, res):
        assert len(a.depth) > 1536, "the case must be long enough for k_chain_big"
        assert a.status == b.status and a.count_chains == b.count_chains
        for name in ("depth", "size", "totalsize", "interval", "next"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert np.array_equal(a.positions, b.positions) and np.array_equal(a.rotations, b.rotations)
    s0 = batch_sets(batch, 0, 1)[0]
    compare_with_oracle(res[0], oracle_run(s0), s0, "long block list")


@pytest.mark.parametrize("seed0", [0, 40, 80])
def test_gpu_randomized_sweep_all_paths(gpu_finder, seed0):
    """seeded families (variants, random, binary, IUPAC, many sequences, shuffled blocks, ragged lengths; lengths 8 to
    6000) through the free choice (with the printed counts), the forced word sort, the word sort stopped early and
    the doubling rounds: every result equals the oracle's.  (30 000 cases of this sweep ran clean on the last build.)"""
    try:
        for seed in range(seed0, seed0 + 6):
            rng = random.Random(1000 + seed)
            cases = [gen_case(rng, max_n=rng.choice([500, 1500, 6000]))[1] for _ in range(150)]
            oras = [oracle_run(s) for s in cases]
            for mode in (0, 6, 4, 5, 10, 12):
                gpu_finder.debug_rounds(mode)
                res = gpu_finder.find_rotations_batch(cases, flags=1 if mode == 0 else 0)
                for i, (r, o, s) in enumerate(zip(res, oras, cases)):
                    compare_with_oracle(r, o, s, f"seed {seed} mode {mode} case {i}")
    finally:
        gpu_finder.debug_rounds(0)


@pytest.mark.parametrize("m,n,nsets,seed,snp", [
    (100, 500, 2, 2, 0.01), (200, 400, 1, 3, 0.01), (150, 300, 3, 12, 0.02),   # short, nearly identical: ties left to the doubling rounds
    (33, 3000, 3, 1, 0.04), (64, 4000, 2, 5, 0.02), (65, 4000, 2, 6, 0.02), (100, 2500, 2, 2, 0.04), (128, 3000, 1, 7, 0.03),
    (200, 2000, 1, 3, 0.04), (256, 5000, 2, 8, 0.02), (257, 3000, 1, 9, 0.03), (300, 2000, 1, 10, 0.04), (256, 16500, 1, 11, 0.01)])
def test_gpu_sets_of_hundreds_of_sequences(gpu_finder, m, n, nsets, seed, snp):
    """sets of 33 .. 300 near-identical sequences: the carried word sort with a CTA per root (k_cywalk_cta), roots ordered
    word by word (k_wsort_words, both CTA sizes), blocks straight from the LCP array up to 256 sequences (k_blockfind2's
    cooperative path) and through the cover array beyond -- free choice, forced carried sort and forced plain word sort
    leave the suffix array, LCP array and rotations of the doubling rounds; the smaller cases are checked against the oracle"""
    from csa_b200.workloads import make_batch
    b = make_batch(nsets=nsets, m=m, n=n, snp=snp, indel=snp / 10, seed=seed, population=True)
    out = {}
    try:
        for mode in (5, 0, 10, 6):
            gpu_finder.debug_rounds(mode)
            res = gpu_finder.find_rotations_batch(b)
            out[mode] = (res,) + tuple(gpu_finder.suffix_array())
    finally:
        gpu_finder.debug_rounds(0)
    for mode in (0, 10, 6):
        assert np.array_equal(out[5][1], out[mode][1]) and np.array_equal(out[5][2], out[mode][2]), mode
        for x, y in zip(out[5][0], out[mode][0]):
            assert np.array_equal(x.rotations, y.rotations) and np.array_equal(x.positions, y.positions), mode
    if m * n < 400000:
        for i, (r, seqs) in enumerate(zip(out[0][0], batch_sets(b))):
            compare_with_oracle(r, oracle_run(seqs), seqs, f"m={m} set {i}")


def test_gpu_one_process_several_gpus():
    """csa_gpu_multi_*: one process, one host thread per GPU, bucket exchange by peer copies; needs two GPUs
    (skipped on a one-GPU box; tests/test_emu_pipeline.py covers the logic on the CPU single-stepper)"""
    import torch
    from csa_b200.api import MultiRotationFinder, RotationFinder
    from csa_b200.workloads import make_batch, batch_sets
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU")
    mf = MultiRotationFinder(min(n, 4))
    one = RotationFinder(device=0)
    try:
        for batch in (make_batch(1, 6, 400_000, 0.01, 0.001, seed=31, population=True), workload_batch("mammals", 5, seed=32)):
            sets = batch_sets(batch)
            ref = one.find_rotations_batch(sets)
            rot, info = mf.find_rotations_batch(sets)
            (depth, size, total, interval, nxt), pos = mf.first.blocks()
            b0 = 0
            for k, r in enumerate(ref):
                q0, q1 = int(batch.set_start[k]), int(batch.set_start[k + 1])
                assert info[k].status == r.status == 0
                assert np.array_equal(rot[q0:q1], r.rotations)
                nb = info[k].nblocks
                assert np.array_equal(depth[b0:b0 + nb], r.depth) and np.array_equal(size[b0:b0 + nb], r.size)
                assert np.array_equal(nxt[b0:b0 + nb], r.next)
                b0 += nb
    finally:
        mf.close()
        one.close()


@pytest.mark.parametrize("name,nsets", [("bacterial", 1), ("mammals", 480), ("sets32", 192)])
def test_gpu_full_size_paths_agree(gpu_finder, name, nsets):
    """BASELINE.json's full sizes (80-100 M suffixes per batch), where the oracle is out of reach: the two independent
    ways to the suffix array -- rank doubling + LCP kernels, word sort with its own LCPs -- must give the same
    rotations, suffix array, LCP array and block lists (they share only the first sort); so must the word sort with a
    column's order carried over to the next (10; what the free choice takes on the bacterial set) and without (11)"""
    batch = workload_batch(name, nsets, seed=1000)
    out = {}
    try:
        for mode in (5, 6, 10, 11):
            gpu_finder.debug_rounds(mode)
            gpu_finder.upload(batch)
            gpu_finder.run()
            rot, info = gpu_finder.download()
            sa, lcp = gpu_finder.suffix_array()
            (depth, size, total, interval, nxt), pos = gpu_finder.blocks()
            out[mode] = (rot, sa, lcp[1:], depth, size, total, interval, nxt, pos, np.array([i.status for i in info]))
    finally:
        gpu_finder.debug_rounds(0)
    assert (out[5][9] == 0).all()
    for mode in (6, 10, 11):
        for x, y in zip(out[5], out[mode]):
            assert np.array_equal(x, y), mode


@pytest.mark.parametrize("name", ["variants256", "bacterial"])
def test_gpu_full_size_against_the_oracle(gpu_finder, name):
    """BASELINE.json configs[2] and configs[4] at their own size against tests/golden/fullsize.json: ONE run of the
    oracle on exactly these bytes (tests/golden/make_fullsize.py; 80 M suffixes for the bacterial set): counts,
    rotations, sha256 of the whole sorted block list and of the suffix array and LCP array.  For the bacterial set
    also through the bucket-sharded entry points (csa_gpu_shard_*, one rank)."""
    import hashlib, json, os
    from common import GOLDEN
    from csa_b200.api import Batch
    fx = json.load(open(os.path.join(GOLDEN, "fullsize.json")))[name]
    batch = workload_batch(name, 1, seed=fx["seed"])
    assert batch.nbases == fx["nbases"]
    sha = lambda x: hashlib.sha256(np.ascontiguousarray(x, dtype=np.int32).tobytes()).hexdigest()

    def check(what):
        rot, info = gpu_finder.download()
        assert info[0].status == fx["status"] == 0, what
        assert [info[0].count_collected, info[0].count_suffixfree, info[0].count_unique, info[0].count_chains] == fx["counts"], what
        assert [int(x) for x in rot] == fx["rotations"], what
        (depth, size, total, interval, nxt), pos = gpu_finder.blocks()
        got = dict(depth=sha(depth), size=sha(size), totalsize=sha(total), interval=sha(interval), next=sha(nxt), positions=sha(pos))
        sa, lcp = gpu_finder.suffix_array()
        got["sa"], got["lcp"] = sha(sa.astype(np.int64)), sha(lcp)
        for k, v in got.items():
            assert v == fx["sha256"][k], (what, k)

    gpu_finder.upload(batch)
    gpu_finder.run(flags=1)
    check("one GPU")
    if name == "bacterial":
        from csa_b200.shard import run_bucket_sharded
        gpu_finder.upload(batch)
        run_bucket_sharded(gpu_finder, 0, 1, cuda=True, flags=1)
        check("bucket entry points, one rank")


def gen_hundreds(rng):
    """a set of 33 .. 300 variants of one ancestor: ACGT, a two-letter alphabet (groups of thousands) or with IUPAC letters
    (the word sort's mask plane), equal or ragged lengths, randomly rotated, identical rotations dropped"""
    from common import drop_rotation_duplicates, mutate
    kind = rng.choice(["acgt", "acgt", "binary", "iupac", "ragged"])
    alphabet = "AC" if kind == "binary" else "ACGT"
    m = rng.choice([rng.randint(33, 70), rng.randint(70, 140), rng.randint(140, 300)])
    n = rng.choice([rng.randint(20, 120), rng.randint(120, 600), rng.randint(600, 2500)])
    base = [rng.choice(alphabet) for _ in range(n)]
    snp, indel = rng.choice([0.002, 0.01, 0.03, 0.1]), rng.choice([0.0, 0.002, 0.01])
    out = []
    for _ in range(m):
        s = mutate(rng, base[:rng.randint(max(2, n // 2), n)] if kind == "ragged" else base, snp, indel, alphabet)
        if kind == "iupac":
            for _ in range(rng.randint(0, 3)):
                if s:
                    s[rng.randrange(len(s))] = rng.choice("NRYKM")
        if len(s) < 2:
            s = s + ["A", "C"]
        r = rng.randrange(len(s))
        out.append("".join(s[r:] + s[:r]).encode())
    out = drop_rotation_duplicates(out)
    return out if len(out) >= 2 else gen_hundreds(rng)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_gpu_randomized_sweep_sets_of_hundreds(gpu_finder, seed):
    """seeded sets of 33 .. 300 sequences, batches of 24 sets of mixed sizes, through the free choice, the forced carried word
    sort (packed and unpacked group table) and the forced plain word sort: every result equals the oracle's.
    (40 seeds of this sweep -- 960 sets, four modes each -- ran clean on the last build.)"""
    rng = random.Random(7000 + seed)
    cases = [gen_hundreds(rng) for _ in range(24)]
    oras = [oracle_run(s) for s in cases]
    try:
        for mode in (0, 10, 12, 6):
            gpu_finder.debug_rounds(mode)
            res = gpu_finder.find_rotations_batch(cases, flags=1 if mode == 0 else 0)
            for i, (r, o, s) in enumerate(zip(res, oras, cases)):
                compare_with_oracle(r, o, s, f"seed {seed} mode {mode} case {i} ({len(s)} sequences)")
    finally:
        gpu_finder.debug_rounds(0)
