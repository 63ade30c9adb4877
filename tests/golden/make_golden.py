#!/usr/bin/env python3
"""Generate tests/golden/golden.json.gz and golden_edge.json.gz from the UNMODIFIED reference (oracle/_ref/CSA_ref).

Run in the build container, where /root/reference exists:
    make -C oracle && python tests/golden/make_golden.py
For every case the reference binary is run as `CSA_ref R in.fa` in a scratch directory.

golden.json.gz -- cases the reference ANSWERS (exit 0).  Kept: the input, the four counts it prints
(csamsa.c:332,338,348,354), the rotations written to in-Rotated.fasta (csamsa.c:421), the text of in-Blocks.csv
(csamsa.c:361) and the sha256 of all five files it writes (-Rotated.fasta, -Blocks.csv, -positions.txt,
-imagemap.txt, -Blocks.bmp).  Cases: the reference's own examples Manual/Primates.txt and Manual/Mammals.txt
(BASELINE.json configs[0], configs[1]); tests/golden/regressions/*.fa (the inputs on which round 1 differed from
the reference, the advisor's examples); seeded synthetic sets of every family of tests/common.py::gen_case,
among them sequences that are powers w^c, sets in which one sequence lies wholly inside all others, and blocks
that hold letters outside ACGT.

golden_edge.json.gz -- cases on which the reference does NOT finish: it is killed by a signal ("dies") or still
runs after 5 s on a few hundred letters ("hangs").  Kept: the input, what it had flushed to stdout, and the
sha256 of -Rotated.fasta when it got that far (then it died in blockLabel, nodeslinkedlists.c:150, on a chain
that closes into a ring).  The tests require the matching classification from oracle and CUDA path.

A case whose input holds a sequence that is an identical rotation of an earlier one AND on which the reference's
answer is bent by the marks the discarded sequence left in its tree (DESIGN.md "identical rotations") is kept in
the form the reference gives on the input WITHOUT the discarded sequences (field "note").
"""
import gzip, hashlib, json, os, random, re, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import gen_case, drop_rotation_duplicates  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "CSA_ref")
FILES = ["-Rotated.fasta", "-Blocks.csv", "-positions.txt", "-imagemap.txt", "-Blocks.bmp"]


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
    """-> dict(outcome="ok"|"dies"|"hangs", ...)"""
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
            rc, out = p.returncode, p.stdout
        except subprocess.TimeoutExpired as e:
            rc, out = "timeout", e.stdout or b""
        out = out.decode("latin1")
        rot_path = os.path.join(tmp, "in-Rotated.fasta")
        rot_bytes = open(rot_path, "rb").read() if os.path.exists(rot_path) else None
        if rc != 0:
            return dict(outcome="hangs" if rc == "timeout" else "dies", stdout_flushed=out[out.find("> Collecting"):] if "> Collecting" in out else "",
                        rotated_sha256=hashlib.sha256(rot_bytes).hexdigest() if rot_bytes is not None else None)
        counts = [int(x) for x in re.findall(r"(\d+) (?:nodes found|nodes left|chains found)", out)]
        if len(counts) != 4 or rot_bytes is None:
            return dict(outcome="message", stdout=out[out.find("> Collecting"):])
        rots = [int(x) for x in re.findall(r"^>.* @ (\d+)$", rot_bytes.decode("latin1"), flags=re.M)]
        blocks = open(os.path.join(tmp, "in-Blocks.csv")).read()
        sha = {sfx: hashlib.sha256(open(os.path.join(tmp, "in" + sfx), "rb").read()).hexdigest() for sfx in FILES}
        return dict(outcome="ok", counts=counts, rotations=rots, blocks_csv=blocks, rotated_sha256=sha["-Rotated.fasta"],
                    files_sha256=sha, stdout=out[out.find("> Collecting"):])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_case(name, descs, seqs, timeout=5):
    """what the reference says on the input; on the input without the discarded sequences when their marks bend it"""
    r = run_ref(descs, seqs, timeout)
    kept = drop_rotation_duplicates([s.encode() for s in seqs])
    if len(kept) != len(seqs):
        kd = [d for d, s in zip(descs, seqs) if s.encode() in kept]
        ks = [s.decode() for s in kept]
        if len(ks) < 2:
            return None
        r2 = run_ref(kd, ks, timeout)
        same = r["outcome"] == r2["outcome"] and all(r.get(k) == r2.get(k) for k in ("counts", "rotations", "blocks_csv", "stdout_flushed"))
        if not same:
            r = r2
            r["note"] = "reference run without the sequences it discards: on the full input the marks they leave in its tree bend the answer"
        descs, seqs = kd, ks
    return dict(name=name, descs=descs, seqs=seqs, **r)


def main():
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle")
    ok, edge = [], []

    def add(c):
        if c is None or c["outcome"] == "message":
            return False
        (ok if c["outcome"] == "ok" else edge).append(c)
        return True

    for name in ("Primates", "Mammals"):
        descs, seqs = read_fasta(f"/root/reference/Manual/{name}.txt")
        c = reference_case(name, descs, seqs, 120)
        assert c["outcome"] == "ok", name
        add(c)
        print(name, c["counts"], c["rotations"])
    for fn in sorted(os.listdir(os.path.join(HERE, "regressions"))):
        descs, seqs = read_fasta(os.path.join(HERE, "regressions", fn))
        c = reference_case("regression_" + fn[:-3], descs, seqs)
        print(fn, c["outcome"], c.get("counts"), c.get("note", "")[:40])
        add(c)
    rng = random.Random(20261018)
    n_ok = n_edge = i = 0
    while n_ok < 110 or n_edge < 40:
        kind, seqs = gen_case(rng, max_n=1500)
        seqs = [s.decode() for s in seqs]
        descs = [f"seq{k}" for k in range(len(seqs))]
        i += 1
        c = reference_case(f"syn{i}_{kind}", descs, seqs)
        if c is None or c["outcome"] == "message":
            continue
        if c["outcome"] == "ok" and n_ok < 110:
            n_ok += add(c)
        elif c["outcome"] != "ok" and n_edge < 40:
            n_edge += add(c)
    for name, cases in (("golden.json.gz", ok), ("golden_edge.json.gz", edge)):
        with gzip.open(os.path.join(HERE, name), "wt", compresslevel=9) as f:
            json.dump(cases, f)
        print(len(cases), "cases ->", name, os.path.getsize(os.path.join(HERE, name)), "bytes")
    kinds = {}
    for c in ok + edge:
        k = (c["name"].split("_")[-1], c["outcome"])
        kinds[k] = kinds.get(k, 0) + 1
    print(sorted(kinds.items()))


if __name__ == "__main__":
    main()
