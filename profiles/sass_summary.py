#!/usr/bin/env python3
"""Per-kernel SASS instruction mix of the built library (cuobjdump -sass) beside ptxas -v's registers and shared memory:
   python profiles/sass_summary.py > profiles/r2z_sass_summary.txt     (after `make -C csa_b200/csrc`)"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "csa_b200", "csrc", "libcsa_gpu.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, stats = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); stats[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        c = stats[cur]
        c["total"] += 1
        c[op.split(".")[0]] += 1
        for pre in ("LDG", "STG"):
            if op.startswith(pre):
                c[pre + (".128" if ".128" in op else ".64" if ".64" in op else ".32")] += 1
names = dict(zip(stats, subprocess.run(["cu++filt"] + list(stats), capture_output=True, text=True).stdout.split("\n")))
regs, cur = {}, None
for line in open(os.path.join(ROOT, "csa_b200", "csrc", "ptxas.log")):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
    m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
    if m and cur:
        regs[cur] = (m.group(1), m.group(3) or "0")
print("# cuobjdump -sass csa_b200/csrc/libcsa_gpu.so (sm_100a).  Per kernel: SASS instructions, registers and static shared memory (ptxas -v),")
print("# global loads and stores by width, shared-memory loads/stores/atomics, global atomics and reductions, warp-level instructions")
print("# (MATCH, VOTE, SHFL, REDUX) and barriers.  No tensor-core (HMMA/UTC*MMA), TMA (UBLKCP/UTMA*) or cluster instruction appears: the path")
print("# is integer and byte work bound by HBM, by L2 latency or by instruction issue (DESIGN.md says which kernel by which).")
cols = ["LDG.32", "LDG.64", "LDG.128", "STG.32", "STG.64", "STG.128", "LDS", "STS", "ATOMS", "ATOMG", "REDG", "MATCH", "VOTE", "SHFL", "REDUX", "BAR"]
print(f"{'kernel':44s} {'inst':>6s} {'regs':>4s} {'smem':>6s} " + " ".join(f"{c:>7s}" for c in cols))
tot = collections.Counter()
for k, c in stats.items():
    r = regs.get(k, ("?", "?"))
    nm = re.sub(r"^void ", "", names[k]).split("(")[0]
    print(f"{nm[:44]:44s} {c['total']:6d} {r[0]:>4s} {r[1]:>6s} " + " ".join(f"{c[x]:7d}" for x in cols))
    tot.update(c)
print("# opcodes seen anywhere:", " ".join(sorted(k for k in tot if k.isupper() and "." not in k)))
