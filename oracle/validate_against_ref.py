#!/usr/bin/env python3
"""Pin the oracle against the UNMODIFIED reference (oracle/_ref/CSA_ref, built by oracle/Makefile).

TEST INFRASTRUCTURE.  For every case both programs get the same multi-FASTA in separate
scratch directories; stdout, <base>-Rotated.fasta and <base>-Blocks.csv must be byte-identical.
Cases: the reference's own Manual/*.txt examples (when /root/reference is present) plus seeded
synthetic sets (mutated+rotated variants, unrelated random sequences, 2-letter alphabets,
IUPAC letters, up to 64 sequences).

  python oracle/validate_against_ref.py --cases 2000 --seed 1
"""
import argparse, os, random, shutil, subprocess, sys, tempfile

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


def gen_case(rng):
    kind = rng.choice(["variants", "variants", "variants", "random", "binary", "iupac", "many", "blocks"])
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
    except subprocess.TimeoutExpired:
        return "timeout", b""


def compare(fasta, timeout=5, keep=None):
    """returns (verdict, detail); verdict in same / differ / ref_failed / oracle_refused"""
    tmp = tempfile.mkdtemp(prefix="csa_val_")
    try:
        res = {}
        for tag, binary in (("ref", REF), ("ora", ORA)):
            d = os.path.join(tmp, tag)
            os.mkdir(d)
            shutil.copy(fasta, os.path.join(d, "in.fa"))
            res[tag] = run(binary, d, "in.fa", timeout)
        (rc_r, out_r), (rc_o, out_o) = res["ref"], res["ora"]
        if rc_o in (3, 4):
            # the oracle declares the input outside the reference's defined behaviour:
            # the reference must indeed crash / hang / or at least not be trusted there
            return "oracle_refused", f"oracle rc={rc_o} ref rc={rc_r}"
        if rc_r != 0:
            # the reference crashes in blockLabel (nodeslinkedlists.c:161, heap overflow) when the
            # chosen chain closes into a cycle -- AFTER it has written -Rotated.fasta (csamsa.c:611)
            a = os.path.join(tmp, "ref", "in-Rotated.fasta")
            b = os.path.join(tmp, "ora", "in-Rotated.fasta")
            if os.path.exists(a) and os.path.exists(b):
                if open(a, "rb").read() == open(b, "rb").read():
                    return "same_rotations_ref_crashed_later", f"ref rc={rc_r}"
                return "differ", "-Rotated.fasta (ref crashed later)"
            return "ref_failed", f"ref rc={rc_r} oracle rc={rc_o}"
        if out_r != out_o:
            return "differ", "stdout"
        for suffix in ("-Rotated.fasta", "-Blocks.csv"):
            a = os.path.join(tmp, "ref", "in" + suffix)
            b = os.path.join(tmp, "ora", "in" + suffix)
            if os.path.exists(a) != os.path.exists(b):
                return "differ", suffix + " existence"
            if os.path.exists(a) and open(a, "rb").read() != open(b, "rb").read():
                return "differ", suffix
        return "same", ""
    finally:
        if keep:
            shutil.copytree(tmp, keep, dirs_exist_ok=True)
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=500)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--keep-failures", default="/tmp/csa_val_fail")
    a = ap.parse_args()
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
        if v == "oracle_refused":
            print("case", i, kind, v, d)
        if v in ("differ", "ref_failed"):
            bad += v == "differ"
            keep = f"{a.keep_failures}/{a.seed}_{i}_{kind}_{v}"
            os.makedirs(keep, exist_ok=True)
            shutil.copy(fa, keep + "/in.fa")
            compare(fa, keep=keep)
            print("case", i, kind, v, d, "->", keep)
    shutil.rmtree(tmp, ignore_errors=True)
    for k in sorted(tally):
        print(k, tally[k])
    print("DIFFERENCES:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
