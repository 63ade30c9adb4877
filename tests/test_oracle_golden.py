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


def test_golden_has_reference_examples(golden):
    names = [c["name"] for c in golden]
    assert "Primates" in names and "Mammals" in names
    prim = golden[names.index("Primates")]
    assert len(prim["seqs"]) == 16 and prim["counts"] == [3004, 2209, 58, 19]
