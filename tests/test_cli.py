"""The drop-in command line `CSA R <multi-fasta>` (csa_b200/host/csa_main.c): stdout, <base>-Rotated.fasta
and <base>-Blocks.csv against the reference's golden vectors and, byte for byte, against the oracle's
CLI.  CPU: the host code linked with the kernel single-stepper (tests/emu/CSA_emu).  -m gpu: the
product binary csa_b200/host/CSA on cuda:0."""
import hashlib
import os
import re
import subprocess
import tempfile

import pytest

from common import EMU_DIR, ORACLE_BIN, ROOT, build_emu, build_oracle


def run_cli(binary, case, tmp, env=None):
    d = tempfile.mkdtemp(dir=tmp)
    with open(os.path.join(d, "in.fa"), "w") as f:
        for desc, s in zip(case["descs"], case["seqs"]):
            f.write(">" + desc + "\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + "\n")
    p = subprocess.run([binary, "R", "in.fa"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120,
                       env=None if env is None else dict(os.environ, **env))
    def rd(name):
        path = os.path.join(d, name)
        return open(path, "rb").read() if os.path.exists(path) else None
    return p.returncode, p.stdout, rd("in-Rotated.fasta"), rd("in-Blocks.csv")


def check_cases(binary, cases, tmp):
    build_oracle()
    for case in cases:
        rc, out, rot, blocks = run_cli(binary, case, tmp)
        assert rc == 0, (case["name"], out[-300:])
        counts = [int(x) for x in re.findall(rb"(\d+) (?:nodes found|nodes left|chains found)", out)]
        assert counts == case["counts"], case["name"]
        assert hashlib.sha256(rot).hexdigest() == case["rotated_sha256"], case["name"]
        assert blocks.decode() == case["blocks_csv"], case["name"]
        rc2, out2, rot2, blocks2 = run_cli(ORACLE_BIN, case, tmp)
        assert (rc2, out2, rot2, blocks2) == (rc, out, rot, blocks), f"{case['name']}: differs from the oracle CLI"


def test_cli_emu(golden, tmp_path):
    build_emu()
    check_cases(os.path.join(EMU_DIR, "CSA_emu"), golden[:2] + golden[2::6], str(tmp_path))


def test_cli_emu_several_gpus(golden, tmp_path):
    """CSA_GPUS=3: one process, three contexts (csa_gpu_multi_*), buckets exchanged by copies; output files and
    stdout equal the one-GPU run's byte for byte"""
    build_emu()
    for case in golden[:2] + golden[3::9]:
        one = run_cli(os.path.join(EMU_DIR, "CSA_emu"), case, str(tmp_path))
        three = run_cli(os.path.join(EMU_DIR, "CSA_emu"), case, str(tmp_path), env={"CSA_GPUS": "3"})
        assert one == three, case["name"]


def test_cli_drops_rotation_duplicates(tmp_path):
    build_emu()
    build_oracle()
    a = "ACGTTGCAAGGCTTAACCGGTATATCCGAGAGTTTACGCA"
    b = "ACGTTGCTAGGCTTAACCGGTATATCGGAGAGTTTACGCA"
    case = dict(descs=["a", "a rotated", "b", "lower"], seqs=[a, a[7:] + a[:7], b, (b[3:] + b[:3]).lower()], name="dups")
    r1 = run_cli(os.path.join(EMU_DIR, "CSA_emu"), case, str(tmp_path))
    r2 = run_cli(ORACLE_BIN, case, str(tmp_path))
    assert r1 == r2 and b"Discarding seq. 2" in r1[1] and b"Discarding seq. 4" in r1[1]


@pytest.mark.gpu
def test_cli_gpu(golden, tmp_path):
    binary = os.path.join(ROOT, "csa_b200", "host", "CSA")
    assert os.path.exists(binary), "csa_b200/host/CSA not built (make -C csa_b200/host)"
    check_cases(binary, golden[:2] + golden[2::4], str(tmp_path))
