#!/usr/bin/env python3
"""tests/golden/fullsize.json: the ORACLE's answer at BASELINE.json's full sizes for the two configs whose reference
run is out of reach (configs[2]: 256 sequences, the reference stops at 64, csamsa.c:22; configs[4]: 16 x 5 Mb, days of
Ukkonen on one core).  One run of oracle/csa_oracle.c per workload, on exactly the bytes bench.py and the -m gpu tests
generate (csa_b200/workloads.py, seed 1000): rotations, counts, and sha256 of depth / size / totalsize / next /
positions of the whole sorted block list and of the suffix array and LCP array.

    python tests/golden/make_fullsize.py            (about an hour of one core and 12 GB for the bacterial set)
"""
import ctypes as C, hashlib, json, os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import OracleResult, oracle_lib  # noqa: E402
from csa_b200.workloads import batch_sets, workload_batch  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()


def run(name, seed=1000):
    seqs = batch_sets(workload_batch(name, 1, seed=seed))[0]
    lib = oracle_lib()
    lib.csa_oracle_run_sa.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.POINTER(OracleResult),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int)]
    m, n = len(seqs), sum(len(s) for s in seqs)
    texts = (C.c_char_p * m)(*seqs)
    sizes = (C.c_int * m)(*[len(s) for s in seqs])
    sa, lcp = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    r = OracleResult()
    t0 = time.time()
    lib.csa_oracle_run_sa(m, texts, sizes, 2**31 - 1, C.byref(r), sa.ctypes.data_as(C.POINTER(C.c_int)), lcp.ctypes.data_as(C.POINTER(C.c_int)))
    nb = r.nblocks
    arr = lambda p, k: np.ctypeslib.as_array(p, shape=(k,)).copy()
    out = dict(workload=name, seed=seed, nseqs=m, nbases=n, oracle_seconds=round(time.time() - t0, 1), status=r.status,
               counts=[r.count_collected, r.count_suffixfree, r.count_unique, r.count_chains], nblocks=nb,
               rotations=[int(r.rotations[k]) for k in range(m)],
               sha256={k: sha(arr(getattr(r, k), nb)) for k in ("depth", "size", "totalsize", "interval", "next")})
    out["sha256"]["positions"] = sha(arr(r.positions, nb * m))
    out["sha256"]["sa"] = sha(sa)
    out["sha256"]["lcp"] = sha(lcp)
    lib.csa_oracle_free(C.byref(r))
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or ["variants256", "bacterial"]
    path = os.path.join(HERE, "fullsize.json")
    have = json.load(open(path)) if os.path.exists(path) else {}
    for name in names:
        have[name] = run(name)
        print(json.dumps(have[name])[:400], flush=True)
        json.dump(have, open(path, "w"), indent=1)
