"""Shared test helpers: the oracle behind ctypes, seeded case generators, result comparison.

TEST INFRASTRUCTURE: this is one of the few places allowed to load oracle/_build/libcsa_oracle.so.
"""
import ctypes as C
import os
import random
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "libcsa_oracle.so")
ORACLE_BIN = os.path.join(ORACLE_DIR, "_build", "csa_oracle")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libcsa_emu.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")
INT_MAX = 2**31 - 1


class OracleResult(C.Structure):
    _fields_ = [("status", C.c_int), ("m", C.c_int), ("count_collected", C.c_int), ("count_suffixfree", C.c_int),
                ("count_unique", C.c_int), ("count_chains", C.c_int), ("nblocks", C.c_int),
                ("depth", C.POINTER(C.c_int)), ("size", C.POINTER(C.c_int)), ("totalsize", C.POINTER(C.c_int)),
                ("interval", C.POINTER(C.c_int)), ("next", C.POINTER(C.c_int)), ("positions", C.POINTER(C.c_int)),
                ("rotations", C.POINTER(C.c_int)), ("letters", C.POINTER(C.c_char_p))]


def build_oracle():
    if not (os.path.exists(ORACLE_LIB) and os.path.exists(ORACLE_BIN)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "_build/csa_oracle", "_build/libcsa_oracle.so"],
                              stdout=subprocess.DEVNULL)


def build_emu():
    subprocess.check_call(["make", "-C", EMU_DIR], stdout=subprocess.DEVNULL)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(ORACLE_LIB)
        lib.csa_oracle_run.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.POINTER(OracleResult)]
        lib.csa_oracle_free.argtypes = [C.POINTER(OracleResult)]
        lib.csa_oracle_gsa.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _oracle = lib
    return _oracle


def oracle_run(seqs, max_interval=INT_MAX):
    """seqs: list of bytes.  Returns a dict of plain numpy arrays."""
    lib = oracle_lib()
    m = len(seqs)
    texts = (C.c_char_p * m)(*seqs)
    sizes = (C.c_int * m)(*[len(s) for s in seqs])
    r = OracleResult()
    lib.csa_oracle_run(m, texts, sizes, max_interval, C.byref(r))
    nb = r.nblocks
    def arr(p, n):
        return np.array([p[i] for i in range(n)], dtype=np.int32) if n and p else np.zeros(0, dtype=np.int32)
    out = dict(status=r.status, count_collected=r.count_collected, count_suffixfree=r.count_suffixfree,
               count_unique=r.count_unique, count_chains=r.count_chains, nblocks=nb)
    if r.status == 0:
        out.update(depth=arr(r.depth, nb), size=arr(r.size, nb), totalsize=arr(r.totalsize, nb),
                   interval=arr(r.interval, nb), next=arr(r.next, nb),
                   positions=arr(r.positions, nb * m).reshape(nb, m), rotations=arr(r.rotations, m),
                   letters=[r.letters[i] for i in range(nb)])
    lib.csa_oracle_free(C.byref(r))
    return out


def oracle_gsa(seqs):
    lib = oracle_lib()
    m = len(seqs)
    n = sum(len(s) for s in seqs)
    texts = (C.c_char_p * m)(*seqs)
    sizes = (C.c_int * m)(*[len(s) for s in seqs])
    sa = np.zeros(n, dtype=np.int32)
    lcp = np.zeros(n, dtype=np.int32)
    rc = lib.csa_oracle_gsa(m, texts, sizes, sa.ctypes.data_as(C.POINTER(C.c_int)), lcp.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return sa, lcp


# ---- seeded generators (same families as oracle/validate_against_ref.py) ---------------------------
def mutate(rng, base, snp, indel, alphabet):
    out = []
    for c in base:
        r = rng.random()
        if r < snp:
            out.append(rng.choice(alphabet))
        elif r < snp + indel / 2:
            continue
        elif r < snp + indel:
            out.append(c)
            out.append(rng.choice(alphabet))
        else:
            out.append(c)
    return out


def drop_rotation_duplicates(seqs):
    """what gencycsuffixtrees.c:518-524 does before the path starts (the host layer's job)"""
    def norm(s):
        return bytes(c if c in b"ACGT" else ord("-") for c in s)
    out = []
    for s in seqs:
        ns = norm(s)
        if any(len(t) == len(s) and ns in (norm(t) + norm(t)) for t in out):
            continue
        out.append(s)
    return out


KINDS = ["variants", "variants", "variants", "random", "binary", "iupac", "many", "blocks", "ragged",
         "periodic", "contained", "iupacruns"]


def gen_case(rng, max_n=3000, kinds=KINDS):
    kind = rng.choice(kinds)
    alphabet = "ACGT"
    if kind == "binary":
        alphabet = "AC"
    m = rng.randint(2, 8)
    if kind == "many":
        m = rng.randint(9, 70)
    n = rng.choice([rng.randint(8, 60), rng.randint(60, 400), rng.randint(400, max_n)])
    seqs = []
    if kind == "random":
        for _ in range(m):
            seqs.append([rng.choice(alphabet) for _ in range(max(4, n // 4 + rng.randint(0, 20)))])
    elif kind == "blocks":
        nb = rng.randint(2, 12)
        blocks = [[rng.choice(alphabet) for _ in range(rng.randint(6, 40))] for _ in range(nb)]
        for _ in range(m):
            order = list(range(nb))
            if rng.random() < 0.5:
                i, j = sorted(rng.sample(range(nb + 1), 2))
                order[i:j] = reversed(order[i:j])
            s = []
            for b in order:
                s += blocks[b] + [rng.choice(alphabet) for _ in range(rng.randint(0, 30))]
            seqs.append(s)
    elif kind == "periodic":
        # powers w^c: the identical rotations of one sequence share a leaf (gencycsuffixtrees.c:507-517)
        if rng.random() < 0.5:
            alphabet = rng.choice(["AC", "ACG", "ACGT"])
        w = [rng.choice(alphabet) for _ in range(rng.choice([1, 2, 3, rng.randint(2, 12), rng.randint(4, 80)]))]
        whole = w * rng.randint(2, 5)
        for _ in range(m):
            r = rng.random()
            if r < 0.35:
                seqs.append(list(whole))
            elif r < 0.5:
                seqs.append(list(w * rng.randint(1, 4)))
            elif r < 0.8:
                seqs.append(mutate(rng, whole, rng.choice([0.0, 0.02, 0.1]), rng.choice([0.0, 0.02]), alphabet))
            else:
                seqs.append(mutate(rng, w * rng.randint(1, 3), 0.05, 0.02, alphabet))
    elif kind == "contained":
        # a whole rotation of one sequence inside the others (insertions only): leaves that hold every sequence
        base = [rng.choice(alphabet) for _ in range(min(n, 300))]
        for k in range(m):
            s = list(base)
            if k and rng.random() < 0.85:
                for _ in range(rng.randint(1, 3)):
                    p = rng.randrange(len(s) + 1)
                    s[p:p] = [rng.choice(alphabet) for _ in range(rng.randint(1, 4))]
            elif k and rng.random() < 0.5:
                s = mutate(rng, s, 0.01, 0.0, alphabet)
            seqs.append(s)
        rng.shuffle(seqs)
    elif kind == "ragged":
        base = [rng.choice(alphabet) for _ in range(n)]
        for _ in range(m):
            cut = rng.randint(max(2, n // 3), n)
            seqs.append(mutate(rng, base[:cut], 0.01, 0.002, alphabet))
    else:
        base = [rng.choice(alphabet) for _ in range(n)]
        snp = rng.choice([0.0, 0.002, 0.01, 0.03, 0.1])
        indel = rng.choice([0.0, 0.0, 0.002, 0.01])
        for _ in range(m):
            seqs.append(mutate(rng, base, snp, indel, alphabet))
        if kind == "iupac":
            for s in seqs:
                for _ in range(rng.randint(0, 3)):
                    if s:
                        s[rng.randrange(len(s))] = rng.choice("NRYKM")
        if kind == "iupacruns":
            # the same places hold different ambiguity letters: blocks with a fifth letter inside
            for _ in range(rng.randint(1, 6)):
                p, ln = rng.randrange(n), rng.randint(1, 5)
                for s in seqs:
                    if rng.random() < 0.8:
                        for q in range(p, min(p + ln, len(s))):
                            s[q] = rng.choice("NRYKMSWBDHV")
    out = []
    for s in seqs:
        if len(s) < 2:
            s = s + ["A", "C"]
        r = rng.randrange(len(s))
        s = s[r:] + s[:r]
        out.append("".join(s).encode())
    out = drop_rotation_duplicates(out)
    if len(out) < 2:
        return gen_case(rng, max_n, kinds)
    return kind, out


def compare_with_oracle(res, ora, seqs, where=""):
    """res: csa_b200.api.SetResult, ora: oracle_run() dict.  Bit-exact or AssertionError."""
    assert res.status == ora["status"], f"{where}: status {res.status} != oracle {ora['status']}"
    if ora["status"] in (3, 4, 5):
        return
    assert res.count_unique == ora["count_unique"], f"{where}: count_unique {res.count_unique} != {ora['count_unique']}"
    if res.count_collected >= 0:
        assert res.count_collected == ora["count_collected"], f"{where}: count_collected"
        assert res.count_suffixfree == ora["count_suffixfree"], f"{where}: count_suffixfree"
    if ora["status"] != 0:
        return
    assert res.count_chains == ora["count_chains"], f"{where}: chains {res.count_chains} != {ora['count_chains']}"
    for name in ("depth", "size", "totalsize", "interval", "next"):
        a, b = getattr(res, name), ora[name]
        assert np.array_equal(a, b), f"{where}: blockslist.{name} differs\n got {a}\n exp {b}"
    assert np.array_equal(res.positions, ora["positions"]), f"{where}: block positions differ"
    assert np.array_equal(res.rotations, ora["rotations"]), f"{where}: rotations {res.rotations} != {ora['rotations']}"
    if res.letters is not None:
        assert [bytes(x) for x in res.letters] == [bytes(x) for x in ora["letters"]], f"{where}: block letters differ"
