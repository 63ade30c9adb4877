#!/usr/bin/env python3
"""Generate tests/golden/golden.json.gz from the UNMODIFIED reference (oracle/_ref/CSA_ref).

Run in the build container, where /root/reference exists:
    make -C oracle && python tests/golden/make_golden.py
For every case the reference binary is run as `CSA_ref R in.fa` in a scratch directory; what is
kept: the input sequences, the four counts it prints (csamsa.c:332,338,348,354), the rotations
written to in-Rotated.fasta (csamsa.c:421), the text of in-Blocks.csv (csamsa.c:361) and the
sha256 of in-Rotated.fasta.  Cases: the reference's own examples Manual/Primates.txt and
Manual/Mammals.txt (BASELINE.json configs[0], configs[1]) and seeded synthetic sets of the same
families as oracle/validate_against_ref.py on which the reference exits 0.
"""
import gzip, hashlib, json, os, random, re, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import gen_case  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "CSA_ref")


def read_fasta(path):
    descs, seqs, cur = [], [], None
    for line in open(path, "rb").read().decode("latin1").splitlines():
        if line.startswith(">"):
            descs.append(line[1:])
            cur = []
            seqs.append(cur)
        elif cur is not None:
            cur.append(line.strip().upper())
    return descs, ["".join(s) for s in seqs]


def run_ref(descs, seqs, timeout=120):
    tmp = tempfile.mkdtemp(prefix="csa_golden_")
    try:
        with open(os.path.join(tmp, "in.fa"), "w") as f:
            for d, s in zip(descs, seqs):
                f.write(">" + d + "\n")
                for i in range(0, len(s), 70):
                    f.write(s[i:i + 70] + "\n")
        try:
            p = subprocess.run([REF, "R", "in.fa"], cwd=tmp, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, timeout=timeout)
        except subprocess.TimeoutExpired:
            return None
        if p.returncode != 0:
            return None
        out = p.stdout.decode("latin1")
        counts = [int(x) for x in re.findall(r"(\d+) (?:nodes found|nodes left|chains found)", out)]
        rot_path = os.path.join(tmp, "in-Rotated.fasta")
        if len(counts) != 4 or not os.path.exists(rot_path):
            return None
        rot_bytes = open(rot_path, "rb").read()
        rots = [int(x) for x in re.findall(r"^>.* @ (\d+)$", rot_bytes.decode("latin1"), flags=re.M)]
        blocks = open(os.path.join(tmp, "in-Blocks.csv")).read()
        return dict(counts=counts, rotations=rots, blocks_csv=blocks, rotated_sha256=hashlib.sha256(rot_bytes).hexdigest())
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def degenerate(seqs):
    """A whole rotation of a shortest sequence occurs (circularly) in every other sequence: the
    reference's tree walk is undefined there (csamsa.c:64 runs off a leaf); such inputs give no
    trustworthy vector even when the binary happens to exit 0."""
    norm = lambda s: "".join(c if c in "ACGT" else "-" for c in s)
    ns = [norm(s) for s in seqs]
    nmin = min(len(s) for s in ns)
    for k, s in enumerate(ns):
        if len(s) != nmin:
            continue
        others = [(t + t)[:len(t) + nmin - 1] for j, t in enumerate(ns) if j != k]
        for r in range(nmin):
            rot = s[r:] + s[:r]
            if all(rot in t for t in others):
                return True
    return False


def main():
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle")
    cases = []
    for name in ("Primates", "Mammals"):
        descs, seqs = read_fasta(f"/root/reference/Manual/{name}.txt")
        r = run_ref(descs, seqs)
        assert r is not None, name
        cases.append(dict(name=name, descs=descs, seqs=seqs, **r))
        print(name, r["counts"], r["rotations"])
    rng = random.Random(20261018)
    n_syn = 0
    while n_syn < 80:
        kind, seqs = gen_case(rng, max_n=1500)
        seqs = [s.decode() for s in seqs]
        descs = [f"seq{k}" for k in range(len(seqs))]
        if degenerate(seqs):
            continue
        r = run_ref(descs, seqs, timeout=5)
        if r is None:
            continue
        cases.append(dict(name=f"syn{n_syn}_{kind}", descs=descs, seqs=seqs, **r))
        n_syn += 1
    with gzip.open(os.path.join(HERE, "golden.json.gz"), "wt", compresslevel=9) as f:
        json.dump(cases, f)
    print(len(cases), "cases ->", os.path.getsize(os.path.join(HERE, "golden.json.gz")), "bytes")


if __name__ == "__main__":
    main()
