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


FILES = ["-Rotated.fasta", "-Blocks.csv", "-positions.txt", "-imagemap.txt", "-Blocks.bmp"]
SHIM_EMU = os.path.join(ROOT, "oracle", "_ref", "CSA_gpu_emu")   # built by `make -C oracle shim-emu` where /root/reference is
SHIM_GPU = os.path.join(ROOT, "oracle", "_ref", "CSA_gpu")       # `make -C oracle shim`; travels to the GPU box prebuilt


def run_cli(binary, case, tmp, env=None, all_files=False):
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
    if all_files:
        return p.returncode, p.stdout, {sfx: rd("in" + sfx) for sfx in FILES}
    return p.returncode, p.stdout, rd("in-Rotated.fasta"), rd("in-Blocks.csv")


def check_all_files(binary, cases, tmp):
    """the reference's own main/loader/drawing code around csa_shim.c: all five files of `CSA R`, byte for byte"""
    for case in cases:
        rc, out, files = run_cli(binary, case, tmp, all_files=True)
        assert rc == 0, (case["name"], out[-300:])
        assert out[out.find(b"> Collecting"):].decode("latin1") == case["stdout"], case["name"]
        for sfx in FILES:
            assert files[sfx] is not None, (case["name"], sfx)
            assert hashlib.sha256(files[sfx]).hexdigest() == case["files_sha256"][sfx], (case["name"], sfx)


def check_edge_cases(binary, cases, tmp, label_dies_by_signal=False):
    """where the reference is killed or never returns, the drop-in says which of the two and stops at the same place"""
    seen = {}
    for case in cases:
        rc, out, rot, _ = run_cli(binary, case, tmp)
        seen[rc] = seen.get(rc, 0) + 1
        flushed = case["stdout_flushed"].encode("latin1")
        if rc == 6:  # the reference read freed memory from "Removing suffixes" on: what it printed after that tells nothing
            flushed = flushed[:flushed.find(b"> Removing suffixes... ") + 23]
        assert out[out.find(b"> Collecting"):].startswith(flushed), case["name"]
        if case["rotated_sha256"] is not None:  # the reference died in blockLabel on a chain that is a ring
            assert rc < 0 if label_dies_by_signal else rc == 5, (case["name"], rc)
            assert hashlib.sha256(rot).hexdigest() == case["rotated_sha256"], case["name"]
        else:
            assert rc in ((3, 4) if case["outcome"] == "hangs" else (3, 6)), (case["name"], rc, case["outcome"])
            assert rot is None, case["name"]
    return seen


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


def test_cli_emu_where_the_reference_does_not_finish(golden_edge, tmp_path):
    build_emu()
    seen = check_edge_cases(os.path.join(EMU_DIR, "CSA_emu"), golden_edge, str(tmp_path))
    assert seen.get(5, 0) and seen.get(3, 0) and seen.get(4, 0), seen


@pytest.mark.skipif(not os.path.exists(SHIM_EMU), reason="oracle/_ref/CSA_gpu_emu needs the reference sources to build")
def test_shim_emu_writes_the_reference_s_five_files(golden, golden_edge, tmp_path):
    """csa_b200/host/csa_shim.c linked with the reference's own objects (kernel bodies single-stepped on the CPU):
    stdout from 'Collecting' on, -Rotated.fasta, -Blocks.csv, -positions.txt, -imagemap.txt, -Blocks.bmp equal
    what the unmodified reference wrote (tests/golden/make_golden.py)"""
    build_emu()
    check_all_files(SHIM_EMU, golden[:2] + golden[2:14] + golden[14::5], str(tmp_path))
    check_edge_cases(SHIM_EMU, golden_edge[::3], str(tmp_path), label_dies_by_signal=True)


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
    check_cases(binary, golden[:2] + golden[2::6], str(tmp_path))


@pytest.mark.gpu
def test_cli_gpu_where_the_reference_does_not_finish(golden_edge, tmp_path):
    binary = os.path.join(ROOT, "csa_b200", "host", "CSA")
    seen = check_edge_cases(binary, golden_edge[1::2], str(tmp_path))
    assert seen.get(5, 0) and (seen.get(3, 0) or seen.get(4, 0)), seen


@pytest.mark.gpu
def test_shim_gpu_writes_the_reference_s_five_files(golden, golden_edge, tmp_path):
    """the reference's own program with its hot path on cuda:0 (oracle/_ref/CSA_gpu = reference objects + csa_shim.c +
    libcsa_gpu.so): all five output files of every golden case byte-identical with the unmodified reference's"""
    assert os.path.exists(SHIM_GPU), "oracle/_ref/CSA_gpu not built (make -C oracle shim, where the reference sources are)"
    # (a process and a CUDA context per case, ~2.5 s each: the examples, the regression inputs, every third synthetic set)
    check_all_files(SHIM_GPU, golden[:12] + golden[12::3], str(tmp_path))
    check_edge_cases(SHIM_GPU, golden_edge[::2], str(tmp_path), label_dies_by_signal=True)
