"""The oracle (oracle/csa_oracle.c) against vectors produced by the UNMODIFIED reference binary
(tests/golden/make_golden.py): counts printed, rotations, and the text of <base>-Blocks.csv."""
import hashlib
from types import SimpleNamespace

import numpy as np

from common import oracle_run
from csa_b200 import host


def as_result(o):
    return SimpleNamespace(**o)


def test_oracle_matches_reference_vectors(golden):
    assert len(golden) >= 50
    for case in golden:
        seqs = [s.encode() for s in case["seqs"]]
        o = oracle_run(seqs)
        assert o["status"] == 0, case["name"]
        got = [o["count_collected"], o["count_suffixfree"], o["count_unique"], o["count_chains"]]
        assert got == case["counts"], case["name"]
        assert list(o["rotations"]) == case["rotations"], case["name"]
        assert host.blocks_csv(as_result(o), seqs) == case["blocks_csv"], case["name"]
        rot = host.rotated_fasta(case["descs"], seqs, o["rotations"])
        assert hashlib.sha256(rot).hexdigest() == case["rotated_sha256"], case["name"]


def test_oracle_agrees_where_the_reference_does_not_finish(golden_edge):
    """tests/golden/golden_edge.json.gz: the reference was killed by a signal or never returned.  The oracle must say
    so: a ring in a printed chain when the reference got as far as -Rotated.fasta (it dies in blockLabel,
    nodeslinkedlists.c:150), else walks-off-a-leaf (3), endless block cycle (4) or frees-what-it-stands-on (5)."""
    assert len(golden_edge) >= 30
    seen = {}
    for case in golden_edge:
        seqs = [s.encode() for s in case["seqs"]]
        o = oracle_run(seqs)
        if case["rotated_sha256"] is not None:
            assert o["status"] == 0, case["name"]
            rot = host.rotated_fasta(case["descs"], seqs, o["rotations"])
            assert hashlib.sha256(rot).hexdigest() == case["rotated_sha256"], case["name"]
            res = as_result(o)
            heads = [b for b in range(o["nblocks"]) if o["totalsize"][b] != -1]
            assert any(host.chain_is_ring(res, b) for b in heads), case["name"]
            seen["ring"] = seen.get("ring", 0) + 1
        else:
            assert o["status"] in ((3, 4) if case["outcome"] == "hangs" else (3, 5)), (case["name"], o["status"], case["outcome"])
            seen[o["status"]] = seen.get(o["status"], 0) + 1
        # the counts the reference had flushed before it stopped
        flushed = [int(x) for x in __import__("re").findall(r"(\d+) (?:nodes found|nodes left|chains found)", case["stdout_flushed"])]
        got = [o["count_collected"], o["count_suffixfree"], o["count_unique"], o["count_chains"]]
        assert got[:len(flushed)] == flushed, (case["name"], got, flushed)
    assert seen.get("ring", 0) and seen.get(3, 0) and seen.get(4, 0), seen


def test_golden_has_reference_examples(golden):
    names = [c["name"] for c in golden]
    assert "Primates" in names and "Mammals" in names
    prim = golden[names.index("Primates")]
    assert len(prim["seqs"]) == 16 and prim["counts"] == [3004, 2209, 58, 19]
