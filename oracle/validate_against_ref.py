#!/usr/bin/env python3
"""Pin the oracle against the UNMODIFIED reference (oracle/_ref/CSA_ref, built by oracle/Makefile).

TEST INFRASTRUCTURE.  For every case both programs get the same multi-FASTA in separate scratch
directories.  Nothing is skipped; every case ends in exactly one of these verdicts:

  same            reference exit 0: stdout, <base>-Rotated.fasta, <base>-Blocks.csv byte-identical,
                  oracle exit 0 too
  same_hang       reference still running after the time limit AND the oracle says "does not
                  terminate" (exit 4) after printing the same stdout the reference flushed
  same_crash      reference killed by a signal AND the oracle predicted exactly that (exit 5 = blockLabel
                  walks a chain that closes into a ring, nodeslinkedlists.c:150; exit 3 = tree walk off
                  a leaf); whatever the reference had flushed to stdout is a prefix of the oracle's and
                  a -Rotated.fasta the reference wrote before dying is byte-identical
  dup_same / dup_same_crash / dup_stale
                  inputs with a sequence that is an identical rotation of an earlier one
                  (gencycsuffixtrees.c:518 "Discarding seq."): see DESIGN.md "identical rotations".
                  dup_same = byte-identical as above; dup_stale = the reference's answer is bent by
                  the sequence marks the discarded sequence left in the tree (gencycsuffixtrees.c:503-506 run
                  before :518 discards) -- then the reference run on the input WITHOUT the discarded
                  sequences must be byte-identical with the oracle (files and counts), which is checked
  differ          anything else.  The script exits 1 if there is one.

  python oracle/validate_against_ref.py --cases 1500 --seed 1

--candidate PATH puts another program in the oracle's place: oracle/_ref/CSA_gpu_emu (the reference's own
main, loader and drawing code with csa_b200/host/csa_shim.c in place of buildGeneralizedTree+analyzeTree,
on the CPU single-stepper of the kernel bodies) or oracle/_ref/CSA_gpu (the same on libcsa_gpu.so, GPU box).
With --all-files the -positions.txt, -imagemap.txt and -Blocks.bmp files are compared too.
"""
import argparse, os, random, re, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "CSA_ref")
ORA = os.path.join(HERE, "_build", "csa_oracle")


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


KINDS = ["variants", "variants", "variants", "random", "binary", "iupac", "iupac", "many", "blocks",
         "periodic", "contained", "dups", "iupacruns"]


def gen_case(rng):
    kind = rng.choice(KINDS)
    alphabet = "ACGT"
    if kind == "binary":
        alphabet = "AC"
    m = rng.randint(2, 8)
    if kind == "many":
        m = rng.randint(9, 64)
    n = rng.choice([rng.randint(8, 60), rng.randint(60, 400), rng.randint(400, 3000)])
    seqs = []
    if kind == "random":
        for _ in range(m):
            seqs.append([rng.choice(alphabet) for _ in range(max(4, n // 4 + rng.randint(0, 20)))])
    elif kind == "blocks":
        # shared blocks in shuffled / partly conserved order, random spacers
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
        # some sequences are exact powers w^k (identical rotations of ONE sequence share a leaf,
        # gencycsuffixtrees.c:507-517), the others variants of one period or of the whole
        if rng.random() < 0.5:
            alphabet = rng.choice(["AC", "ACG", "ACGT"])
        w = [rng.choice(alphabet) for _ in range(rng.choice([1, 2, 3, rng.randint(2, 12), rng.randint(4, 80)]))]
        k = rng.randint(2, 5)
        whole = w * k
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
        # a whole rotation of one sequence inside the others: insertions only
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
    else:
        base = [rng.choice(alphabet) for _ in range(n)]
        snp = rng.choice([0.002, 0.01, 0.03, 0.1]) if kind != "dups" else rng.choice([0.0, 0.0, 0.01])
        indel = rng.choice([0.0, 0.0, 0.002, 0.01])
        for _ in range(m):
            seqs.append(mutate(rng, base, snp, indel, alphabet))
        if kind == "dups":
            for _ in range(rng.randint(1, 3)):
                seqs.insert(rng.randrange(len(seqs) + 1), list(rng.choice(seqs)))
            seqs = seqs[:64]
        if kind == "iupac":
            for s in seqs:
                for _ in range(rng.randint(0, 3)):
                    if s:
                        s[rng.randrange(len(s))] = rng.choice("NRYKM")
        if kind == "iupacruns":
            # the same places hold (different) ambiguity letters in several sequences: blocks with
            # a fifth letter inside, whose spelling depends on the edge-creating occurrence
            for _ in range(rng.randint(1, 6)):
                p, ln = rng.randrange(n), rng.randint(1, 5)
                for s in seqs:
                    if rng.random() < 0.8:
                        for q in range(p, min(p + ln, len(s))):
                            s[q] = rng.choice("NRYKMSWBDHV")
    out = []
    for k, s in enumerate(seqs):
        if len(s) < 2:
            s = s + ["A", "C"]
        r = rng.randrange(len(s))
        s = s[r:] + s[:r]
        out.append((f"seq{k} r{r}", "".join(s)))
    return kind, out


def write_fasta(path, seqs, width=70):
    with open(path, "w") as f:
        for d, s in seqs:
            f.write(">" + d + "\n")
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")


def run(binary, d, name, timeout):
    try:
        p = subprocess.run([binary, "R", name], cwd=d, stdin=subprocess.DEVNULL,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
        return p.returncode, p.stdout
    except subprocess.TimeoutExpired as e:
        return "timeout", (e.stdout or b"")


def read(path):
    return open(path, "rb").read() if os.path.exists(path) else None


FILES = ["-Rotated.fasta", "-Blocks.csv"]


def files_equal(da, db, suffixes=None):
    suffixes = suffixes or FILES
    for suffix in suffixes:
        a, b = read(os.path.join(da, "in" + suffix)), read(os.path.join(db, "in" + suffix))
        if a != b:
            return suffix
    return None


def canonical(seq):
    """the reference compares letters with everything outside ACGT folded into one (gencycsuffixtrees.c:283)"""
    return re.sub("[^ACGT]", "-", seq)


def drop_duplicates(seqs):
    """what gencycsuffixtrees.c:518 intends: a sequence that is a rotation of an earlier kept one goes"""
    kept = []
    for d, s in seqs:
        c = canonical(s)
        if any(len(c) == len(canonical(t)) and c in canonical(t) * 2 for _, t in kept):
            continue
        kept.append((d, s))
    return kept


def strip_loader_lines(out):
    """stdout from 'Collecting' on: the counts and the chain list (the loader lines name the sequences)"""
    i = out.find(b"> Collecting")
    return out[i:] if i >= 0 else out


def parse_fasta(path):
    seqs, d, s = [], None, []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if d is not None:
                seqs.append((d, "".join(s)))
            d, s = line[1:], []
        else:
            s.append(line)
    if d is not None:
        seqs.append((d, "".join(s)))
    return seqs


def compare(fasta, timeout=5, keep=None):
    """returns (verdict, detail)"""
    tmp = tempfile.mkdtemp(prefix="csa_val_")
    try:
        res = {}
        for tag, binary in (("ref", REF), ("ora", ORA)):
            d = os.path.join(tmp, tag)
            os.mkdir(d)
            shutil.copy(fasta, os.path.join(d, "in.fa"))
            res[tag] = run(binary, d, "in.fa", timeout if tag == "ref" else 10 * timeout)
        (rc_r, out_r), (rc_o, out_o) = res["ref"], res["ora"]
        dref, dora = os.path.join(tmp, "ref"), os.path.join(tmp, "ora")
        dup = b"Discarding seq." in out_o
        pre = "dup_" if dup else ""

        def verdict():
            if rc_o == 6:  # use-after-free in the reference: whatever it does, it told us nothing
                return (pre + "ref_undefined", f"ref {rc_r}") if out_o.startswith(out_r[:out_r.find(b"> Removing suffixes... ") + 23]) else ("differ", "stdout before the undefined step")
            if rc_r == 0:
                if rc_o != 0:
                    return "differ", f"reference exit 0, oracle exit {rc_o}"
                if out_r != out_o:
                    return "differ", "stdout"
                bad = files_equal(dref, dora)
                return ("differ", bad) if bad else (pre + "same", "")
            # the reference did not finish: the oracle must have said so, and agree with all it left behind
            if not out_o.startswith(out_r):
                return "differ", f"stdout flushed before the reference stopped (ref {rc_r}, oracle {rc_o})"
            rot = read(os.path.join(dref, "in-Rotated.fasta"))
            if rot is not None and rot != read(os.path.join(dora, "in-Rotated.fasta")):
                return "differ", f"-Rotated.fasta (ref {rc_r}, oracle {rc_o})"
            # exit 3: "walks off a leaf": NULL dereference or endless loop, whichever the tree holds
            if rc_r == "timeout":
                return (pre + "same_hang", "") if rc_o in (3, 4) else ("differ", f"reference hangs, oracle exit {rc_o}")
            if rc_o in (3, 5):
                return pre + "same_crash", f"ref {rc_r} oracle {rc_o}"
            # a candidate that runs the reference's own blockLabel dies in it exactly as the reference does
            # (both wrote the same -Rotated.fasta first: csamsa.c:611 comes before :612)
            if isinstance(rc_o, int) and rc_o < 0 and rot is not None:
                return pre + "same_crash", f"ref {rc_r} candidate {rc_o} (both in blockLabel)"
            return "differ", f"reference died ({rc_r}), oracle exit {rc_o}"

        v, d = verdict()
        if v == "differ" and dup:
            # the marks a discarded sequence left in the tree (see the module docstring): the reference on
            # the input without the discarded sequences must agree with the oracle byte for byte
            kept = drop_duplicates(parse_fasta(fasta))
            d2 = os.path.join(tmp, "ref_dedup")
            os.mkdir(d2)
            write_fasta(os.path.join(d2, "in.fa"), kept)
            rc2, out2 = run(REF, d2, "in.fa", timeout)
            ok = False
            if rc2 == 0 and rc_o == 0:
                ok = strip_loader_lines(out2) == strip_loader_lines(out_o) and files_equal(d2, dora, ("-Blocks.csv",)) is None
                if ok:  # same rotations for the kept sequences (the descriptions are the same, the file too)
                    ok = read(os.path.join(d2, "in-Rotated.fasta")) == read(os.path.join(dora, "in-Rotated.fasta"))
            elif rc2 == "timeout":
                ok = rc_o in (3, 4) and strip_loader_lines(out_o).startswith(strip_loader_lines(out2))
            elif rc2 != 0:
                rot2 = read(os.path.join(d2, "in-Rotated.fasta"))
                died_in_label = isinstance(rc_o, int) and rc_o < 0 and rot2 is not None
                ok = (rc_o in (3, 5) or died_in_label) and strip_loader_lines(out_o).startswith(strip_loader_lines(out2)) \
                    and (rot2 is None or rot2 == read(os.path.join(dora, "in-Rotated.fasta")))
            if ok:
                return "dup_stale", f"ref on the full input: {rc_r}; {d}"
            return "differ", f"{d}; and the reference without the discarded sequences (exit {rc2}) disagrees too"
        return v, d
    finally:
        if keep:
            shutil.copytree(tmp, keep, dirs_exist_ok=True)
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    global ORA
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--keep-failures", default="/tmp/csa_val_fail")
    ap.add_argument("--keep-kind", default="", help="also keep the inputs of cases with this verdict")
    ap.add_argument("--candidate", default=os.path.join(HERE, "_build", "csa_oracle"))
    ap.add_argument("--all-files", action="store_true")
    a = ap.parse_args()
    ORA = os.path.abspath(a.candidate)
    if a.all_files:
        FILES.extend(["-positions.txt", "-imagemap.txt", "-Blocks.bmp"])
    if not (os.path.exists(REF) and os.path.exists(ORA)):
        sys.exit("build first: make -C oracle")
    tally = {}
    bad = 0
    manual = "/root/reference/Manual"
    if os.path.isdir(manual):
        for name in ("Primates.txt", "Mammals.txt"):
            v, d = compare(os.path.join(manual, name), timeout=120)
            print(name, v, d)
            bad += v != "same"
    rng = random.Random(a.seed)
    tmp = tempfile.mkdtemp(prefix="csa_gen_")
    for i in range(a.cases):
        kind, seqs = gen_case(rng)
        fa = os.path.join(tmp, "c.fa")
        write_fasta(fa, seqs)
        v, d = compare(fa)
        tally[(kind, v)] = tally.get((kind, v), 0) + 1
        if v == "differ" or v == a.keep_kind:
            bad += v == "differ"
            keep = f"{a.keep_failures}/{a.seed}_{i}_{kind}_{v}"
            os.makedirs(keep, exist_ok=True)
            shutil.copy(fa, keep + "/in.fa")
            print("case", i, kind, v, d, "->", keep, flush=True)
    shutil.rmtree(tmp, ignore_errors=True)
    for k in sorted(tally):
        print(k, tally[k])
    verdicts = {}
    for (k, v), c in tally.items():
        verdicts[v] = verdicts.get(v, 0) + c
    print("TOTAL", {v: verdicts[v] for v in sorted(verdicts)})
    print("DIFFERENCES:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
