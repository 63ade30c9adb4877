"""Host side of the `./CSA R <multi-fasta>` path in Python: what the reference does around
buildGeneralizedTree()/analyzeTree() -- loading (csamsa.c:437 LoadSequences), dropping sequences
that are rotations of an earlier one (gencycsuffixtrees.c:518-524), and the text outputs
(csamsa.c:421 saveRotatedSequences, csamsa.c:361 createImageAndShowResults: <base>-Blocks.csv).
The C program csa_b200/host/csa_main.c is the drop-in CLI; this module serves tests and bench.
"""
from typing import List, Sequence, Tuple

from .api import SetResult

IUPAC = b"ACGTRYSWKMDHBVN"
MAXNUMBEROFSEQS = 64  # csamsa.c:22


def load_sequences(path: str, max_seqs: int = MAXNUMBEROFSEQS) -> Tuple[List[str], List[bytes]]:
    """csamsa.c:437: '>' starts a record, the rest of that line is the description; letters are
    upper-cased, line ends / '-' / blanks skipped; a record with any other character is dropped."""
    data = open(path, "rb").read()
    descs, seqs = [], []
    pos = data.find(b">")
    if pos < 0:
        raise ValueError("No sequences in file")
    n = len(data)
    while pos < n and len(seqs) < max_seqs:
        pos = data.find(b">", pos)
        if pos < 0:
            break
        pos += 1
        e = pos
        while e < n and data[e] not in b"\r\n":
            e += 1
        desc = data[pos:e].decode("latin1")
        pos = min(e + 1, n)
        end = data.find(b">", pos)
        if end < 0:
            end = n
        out = bytearray()
        bad = False
        stop = end
        for i in range(pos, end):
            c = data[i]
            if c in b"\n\r\0- ":
                continue
            if 97 <= c <= 122:
                c -= 32
            if c in IUPAC:
                out.append(c)
            else:
                bad = True
                stop = i
                break
        pos = stop + 1 if bad else end
        if bad or not out:
            continue
        descs.append(desc)
        seqs.append(bytes(out))
    return descs, seqs


def _norm(s: bytes) -> bytes:
    return bytes(c if c in b"ACGT" else 45 for c in s)


def drop_rotation_duplicates(descs: Sequence[str], seqs: Sequence[bytes]):
    """gencycsuffixtrees.c:518-524: a sequence that is a rotation of an earlier one is discarded
    (letters outside ACGT all compare equal).  Returns (descs, seqs, dropped original indices)."""
    kd, ks, dropped = [], [], []
    for i, (d, s) in enumerate(zip(descs, seqs)):
        ns = _norm(s)
        if any(len(t) == len(s) and ns in (_norm(t) * 2) for t in ks):
            dropped.append(i)
            continue
        kd.append(d)
        ks.append(s)
    return kd, ks, dropped


def chain_is_ring(res: SetResult, b: int) -> bool:
    """the chain that starts at block b closes into a ring: blockLabel (nodeslinkedlists.c:150) never returns"""
    cur, steps = b, 0
    while cur != -1:
        steps += 1
        if steps > len(res.depth):
            return True
        cur = int(res.next[cur])
    return False


def block_label(res: SetResult, b: int, seqs: Sequence[bytes]) -> str:
    """nodeslinkedlists.c:128 blockLabel: the blocks of a chain spelled out, gaps as '-' (up to 7)
    or '-(n)-'; a negative gap eats letters back.  The letters of a block are those of res.letters
    (csa_gpu_batch_block_letters: spelled from the text that created each tree edge); without them the
    letters of the block's place in sequence 0, which differ only where a block holds a letter outside ACGT."""
    s0, n0 = seqs[0], len(seqs[0])
    label = bytearray()
    ln = 0
    cur, guard = b, 0
    nb = len(res.depth)
    while cur != -1 and guard <= nb:
        d, p0 = int(res.depth[cur]), int(res.positions[cur][0])
        if getattr(res, "letters", None) is not None:
            piece = bytes(res.letters[cur])
        else:
            piece = bytes(s0[(p0 + i) % n0] for i in range(d))
        label[ln:ln + d] = piece
        ln += d
        g = int(res.interval[cur])
        if g < 0:
            ln = max(0, ln + g)
        elif g > 7:
            t = b"-(%d)-" % g
            label[ln:ln + len(t)] = t
            ln += len(t)
        else:
            label[ln:ln + g] = b"-" * g
            ln += g
        cur = int(res.next[cur])
        guard += 1
    out = bytes(label[:ln])
    z = out.find(b"\0")
    return (out if z < 0 else out[:z]).decode("latin1")


def blocks_csv(res: SetResult, seqs: Sequence[bytes]) -> str:
    m = len(seqs)
    lines = ["Length,Sequence" + "".join(f",Position_{i + 1}" for i in range(m))]
    for b in range(len(res.depth)):
        if res.totalsize[b] == -1:
            continue
        lines.append(f"{int(res.totalsize[b])},{block_label(res, b, seqs)}" + "".join(f",{int(p)}" for p in res.positions[b]))
    return "\n".join(lines) + "\n"


def rotated_fasta(descs: Sequence[str], seqs: Sequence[bytes], rotations) -> bytes:
    out = bytearray()
    for d, s, r in zip(descs, seqs, rotations):
        r = int(r)
        out += b">" + d.encode("latin1") + b" @ %d\n" % r + s[r:] + s[:r] + b"\n"
    return bytes(out)
