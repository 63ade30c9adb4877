"""Synthetic circular-genome workloads of the shapes BASELINE.json names (no network, no data
files on the GPU box): sets of randomly rotated variants of a random ancestor with substitutions
and indels.  Everything is seeded numpy; the same seed gives the same bytes on every box.
"""
import numpy as np

from .api import Batch

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

# lengths of the 12 mitogenomes of the reference's Manual/Mammals.txt (BASELINE.json configs[1])
MAMMALS_LENGTHS = [16651, 17488, 17019, 16338, 16295, 16402, 16660, 16896, 16571, 16727, 16641, 17734]


def variant(rng, base, snp, indel, target_len=None):
    """one descendant of `base` (uint8 letters): substitutions, deletions, insertions, a random
    rotation; trimmed/padded to target_len when given"""
    n = len(base)
    s = base.copy()
    hit = rng.random(n) < snp  # snp: one rate, or one rate per site
    s[hit] = ACGT[rng.integers(0, 4, int(hit.sum()))]
    if indel > 0:
        keep = rng.random(n) >= indel / 2
        s = s[keep]
        ins = np.nonzero(rng.random(len(s)) < indel / 2)[0]
        s = np.insert(s, ins, ACGT[rng.integers(0, 4, len(ins))])
    if target_len is not None:
        if len(s) > target_len:
            cut = int(rng.integers(0, len(s) - target_len + 1))
            s = np.concatenate([s[:cut], s[cut + len(s) - target_len:]])
        elif len(s) < target_len:
            at = int(rng.integers(0, len(s) + 1))
            s = np.concatenate([s[:at], ACGT[rng.integers(0, 4, target_len - len(s))], s[at:]])
    return np.roll(s, -int(rng.integers(0, len(s))))


def make_set(rng, m, n, snp, indel, lengths=None):
    base = ACGT[rng.integers(0, 4, n)]
    if isinstance(snp, tuple):
        # (conserved rate, variable rate, conserved fraction): per-200-base segments, like the
        # rRNA/tRNA genes against the control region and third codon positions of a mitogenome
        lo, hi, frac = snp
        seg = np.where(rng.random((n + 199) // 200) < frac, lo, hi)
        snp = np.repeat(seg, 200)[:n]
    return [variant(rng, base, snp, indel, None if lengths is None else lengths[k]) for k in range(m)]


def population_set(rng, m, n, site_rate, indel_rate):
    """m members of one population: `site_rate` of the sites are polymorphic (substitution, the
    alternative allele at a random frequency), `indel_rate` carry a 1-base insertion or deletion;
    every member is then rotated at random.  Between polymorphic sites all members agree, which
    is what gives a set its common blocks."""
    base = ACGT[rng.integers(0, 4, n)]
    sub_sites = np.nonzero(rng.random(n) < site_rate)[0]
    sub_freq = rng.random(len(sub_sites))
    sub_alt = (base[sub_sites].astype(np.int64) * 0 + ACGT[(np.searchsorted(ACGT, base[sub_sites]) + rng.integers(1, 4, len(sub_sites))) % 4])
    ind_sites = np.nonzero(rng.random(n) < indel_rate)[0]
    ind_freq = rng.random(len(ind_sites))
    ind_is_del = rng.random(len(ind_sites)) < 0.5
    ind_letter = ACGT[rng.integers(0, 4, len(ind_sites))]
    out = []
    for _ in range(m):
        s = base.copy()
        c = rng.random(len(sub_sites)) < sub_freq
        s[sub_sites[c]] = sub_alt[c]
        ci = rng.random(len(ind_sites)) < ind_freq
        keep = np.ones(n, dtype=bool)
        keep[ind_sites[ci & ind_is_del]] = False
        ins_at = ind_sites[ci & ~ind_is_del]
        ins_letters = ind_letter[ci & ~ind_is_del]
        # insert first (positions refer to the ancestor), then delete
        s2 = np.insert(s, ins_at, ins_letters)
        keep2 = np.insert(keep, ins_at, True)
        s2 = s2[keep2]
        out.append(np.roll(s2, -int(rng.integers(0, len(s2)))))
    return out


def make_batch(nsets, m, n, snp, indel, seed, lengths=None, population=False) -> Batch:
    rng = np.random.default_rng(seed)
    seqs = []
    for _ in range(nsets):
        seqs.extend(population_set(rng, m, n, snp, indel) if population else make_set(rng, m, n, snp, indel, lengths))
    text_start = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=text_start[1:])
    set_start = np.arange(0, nsets * m + 1, m, dtype=np.int32)
    return Batch.from_arrays(np.concatenate(seqs).tobytes(), text_start, set_start)


def batch_sets(batch: Batch, lo=0, hi=None):
    """the sets of a Batch as lists of bytes (for the oracle / the reference CLI)"""
    hi = batch.nsets if hi is None else hi
    out = []
    for s in range(lo, hi):
        q0, q1 = int(batch.set_start[s]), int(batch.set_start[s + 1])
        out.append([batch.text[int(batch.text_start[k]):int(batch.text_start[k + 1])] for k in range(q0, q1)])
    return out


WORKLOADS = {
    # name: (sequences per set, ancestor length, substitution rate, indel rate, lengths, what it is)
    "mammals": (12, 16800, (0.004, 0.25, 0.3), 0.01, MAMMALS_LENGTHS,
                "configs[1]: Mammals.txt-shaped set (12 mitogenomes of 16.3-17.7 kb; 30% conserved segments at "
                "0.4% substitutions, the rest at 25%; 1% indels)"),
    # population model (lengths == "population"): 1% of the sites polymorphic, 0.1% with an indel
    "variants256": (256, 16500, 0.01, 0.001, "population",
                    "configs[2]: 256 randomly rotated 16.5 kb mitogenome variants (1% SNP sites, 0.1% indel sites)"),
    "sets32": (32, 16500, 0.01, 0.001, "population",
               "configs[3]: independent mitogenome sets of 32 sequences (16.5 kb, 1% SNP sites, 0.1% indel sites)"),
    "bacterial": (16, 5_000_000, 0.01, 0.001, "population",
                  "configs[4]: 16 circular 5 Mb chromosomes sharing syntenic blocks (1% SNP sites, 0.1% indel sites)"),
}


def workload_batch(name, nsets, seed) -> Batch:
    m, n, snp, indel, lengths, _ = WORKLOADS[name]
    if lengths == "population":
        return make_batch(nsets, m, n, snp, indel, seed, population=True)
    return make_batch(nsets, m, n, snp, indel, seed, lengths)
