// pipeline.cuh -- the rotation-finding path of fjdf/CSA (`./CSA R`) as device kernels.
//
// Reference path replaced (see include/csa_gpu.h and DESIGN.md):
//   buildGeneralizedTree   gencycsuffixtrees.c:418   -> stage 1+2: generalized cyclic suffix array + LCP
//   collectNodes           csamsa.c:64               -> stage 3: LCP intervals holding every sequence
//   removeSuffixNodes      csamsa.c:80               ->          ... that cannot be extended to the left
//   removeNonUniqueNodes   csamsa.c:283              ->          ... of exactly m suffixes
//   insertSortedItem       nodeslinkedlists.c:36     -> stage 4: block order (depth, DFS visiting order)
//   collectNodeChains      csamsa.c:135              -> stage 5: block chaining
//   sortList/getRotations  nodeslinkedlists.c:59, csamsa.c:311 -> stage 5: chain order, cut points
//
// A batch holds many independent sequence sets.  All suffixes (= rotations) of the batch live
// in ONE index space g = seq_off[k] + p; the set number is the most significant part of every
// sort key, so suffixes of different sets never mix and one launch serves the whole batch.
#pragma once
#include "prims.cuh"

#define CSA_K0 13          // letters in the initial sort key
#define CSA_LETTER_BITS 3  // A C G T other -> 0..4
#define CSA_K0_BITS (CSA_K0 * CSA_LETTER_BITS)
#define CSA_NONE 0xFFFFFFFFu

struct BatchView {
    int nsets;
    u32 M;                     // sequences
    u32 N;                     // bases == suffixes
    const u32 *seq_off;        // [M+1] first base of sequence k
    const u32 *seq_set;        // [M]   set of sequence k
    const u32 *set_seq0;       // [nsets+1] first sequence of set s
    const u32 *set_base0;      // [nsets+1] first base (== first SA index) of set s
    const u32 *set_nmin;       // [nsets] shortest sequence of set s
    const u32 *seq_nmin;       // [M] the same by sequence: set_nmin[seq_set[k]]
    const u64 *dbl_off;        // [M+1] first base of sequence k in the doubled, packed text
    u32 *seqof;                // [N] sequence of base g
    unsigned char *code;       // [N] letter codes 0..4
    u64 *p2;                   // doubled text, 2 bits per base, 32 bases per word
    u32 *pm;                   // doubled text, 1 bit per base: letter is not A/C/G/T
};

// gencycsuffixtrees.c:283/297: every letter that is not A/C/G/T is one fifth letter
HD unsigned code_of_letter(unsigned char c) {
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

HD u32 upper_bound_u32(const u32 *a, u32 n, u32 x) { // first index with a[idx] > x
    u32 lo = 0, hi = n;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (LDG(a + mid) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
HD u32 upper_bound_u64(const u64 *a, u32 n, u64 x) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (LDG(a + mid) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// set of SA position i: the suffix array is grouped by set, borders in set_base0 (a small, cached table)
HD u32 set_of_pos(const BatchView &v, u32 i) { return upper_bound_u32(v.set_base0, (u32)v.nsets + 1, i) - 1; }

HD u32 seq_of(const BatchView &v, u32 g) { return LDG(v.seqof + g); }
// The same by a search in the table of sequence starts: for a batch of a few long sequences (bacterial
// chromosomes) the table sits in L1 and the search beats a 4-byte gather from an N-sized array in HBM in the
// kernels that do little else (measured on 16 x 5 Mb: k_blockfind 2.65 -> 1.41 ms, k_colorkey 1.45 -> 0.45 ms);
// for thousands of short sequences the gather wins.
HD u32 seq_of_few(const BatchView &v, u32 g) {
    if (v.M <= 64u) return upper_bound_u32(v.seq_off, v.M + 1, g) - 1;
    return LDG(v.seqof + g);
}

// position h letters further round the circle
HD u32 cyc_add(const BatchView &v, u32 g, u32 h) {
    u32 k = seq_of(v, g);
    u32 off = LDG(v.seq_off + k), n = LDG(v.seq_off + k + 1) - off;
    u32 p = g - off;
    if (h >= n) h %= n;
    u32 q = p + h; // < 2n <= 2^32 is guaranteed by N < 2^31
    if (q >= n) q -= n;
    return off + q;
}

// ---- stage 0: encode + pack ------------------------------------------------------------------
struct EncodeArgs { BatchView v; const unsigned char *raw; u32 *any_other; };
HD void encode_body(long long i, const EncodeArgs &a) {
    u32 g = (u32)i;
    unsigned c = code_of_letter(a.raw[g]);
    a.v.code[g] = (unsigned char)c;
    if (c > 3) *a.any_other = 1u; // same value from every writer
    u32 k = upper_bound_u32(a.v.seq_off, a.v.M + 1, g) - 1;
    a.v.seqof[g] = k;
}
#ifdef CSA_EMU
MAP_KERNEL_N(encode, EncodeArgs, 6)
#else
// 16 bases per thread: one 16-byte load, one search for the sequence of the first base (the others follow
// by comparing with the next border), 16 + 64 bytes stored as five 16-byte words
__global__ void __launch_bounds__(256) k_encode(long long n, EncodeArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long g0 = t * 16;
    if (g0 >= n) return;
    if (g0 + 16 > n) { for (long long g = g0; g < n; g++) encode_body(g, a); return; }
    const uint4 w = *reinterpret_cast<const uint4 *>(a.raw + g0);
    const u32 in[4] = {w.x, w.y, w.z, w.w};
    u32 k = upper_bound_u32(a.v.seq_off, a.v.M + 1, (u32)g0) - 1;
    u32 nextb = LDG(a.v.seq_off + k + 1);
    u32 outc[4], outk[16];
    bool other = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        u32 oc = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const u32 g = (u32)g0 + 4 * q + b;
            while (g >= nextb) { k++; nextb = LDG(a.v.seq_off + k + 1); }
            const unsigned c = code_of_letter((unsigned char)(in[q] >> (8 * b)));
            other |= c > 3;
            oc |= c << (8 * b);
            outk[4 * q + b] = k;
        }
        outc[q] = oc;
    }
    *reinterpret_cast<uint4 *>(a.v.code + g0) = make_uint4(outc[0], outc[1], outc[2], outc[3]);
#pragma unroll
    for (int q = 0; q < 4; q++)
        *reinterpret_cast<uint4 *>(a.v.seqof + g0 + 4 * q) = make_uint4(outk[4 * q], outk[4 * q + 1], outk[4 * q + 2], outk[4 * q + 3]);
    if (other) *a.any_other = 1u;
}
static inline void launch_encode(Exec &ex, long long n, EncodeArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_encode", 6.0 * n);
    const long long threads = (n + 15) / 16;
    k_encode<<<(unsigned)((threads + 255) / 256), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// one thread per 32-base word of the doubled text: sequence k occupies bases
// [dbl_off[k], dbl_off[k+1]) = s_k s_k s_k... so that any window p..p+n+63 reads without wrap
struct PackArgs { BatchView v; };
HD void pack_body(long long w, const PackArgs &a) {
    u64 x0 = (u64)w * 32;
    u32 k = upper_bound_u64(a.v.dbl_off, a.v.M + 1, x0) - 1;
    if (k >= a.v.M) { a.v.p2[w] = 0; a.v.pm[w] = 0; return; } // the two guard words at the end
    u32 off = LDG(a.v.seq_off + k), n = LDG(a.v.seq_off + k + 1) - off;
    u32 p = (u32)((x0 - LDG(a.v.dbl_off + k)) % n);
    u64 w2 = 0;
    u32 wm = 0;
    for (int t = 0; t < 32; t++) {
        unsigned c = a.v.code[off + p];
        if (c > 3) wm |= 1u << t; else w2 |= (u64)c << (2 * t);
        if (++p == n) p = 0;
    }
    a.v.p2[w] = w2;
    a.v.pm[w] = wm;
}
MAP_KERNEL(pack, PackArgs, 44)

// ---- reading the packed, doubled text ------------------------------------------------------------
HD u64 fetch2(const u64 *p2, u64 x) { // 32 bases starting at doubled base x
    u64 j = x >> 5;
    unsigned s = (unsigned)(x & 31) * 2;
    u64 lo = LDG(p2 + j);
    if (s == 0) return lo;
    return (lo >> s) | (LDG(p2 + j + 1) << (64 - s));
}
HD u32 fetchm(const u32 *pm, u64 x) {
    u64 j = x >> 5;
    unsigned s = (unsigned)(x & 31);
    u32 lo = LDG(pm + j);
    if (s == 0) return lo;
    return (lo >> s) | (LDG(pm + j + 1) << (32 - s));
}
HD int ctz64(u64 x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
HD int ctz32(u32 x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// 32 letters as a number that compares like the letters do: first letter in the top bits
HD u64 lexkey2(u64 w) {
#if defined(__CUDA_ARCH__)
    u64 r = __brevll(w);
#else
    u64 r = w;
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
    r = ((r >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((r & 0x0F0F0F0F0F0F0F0Full) << 4);
    r = __builtin_bswap64(r);
#endif
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1); // bit pairs back in order
}

// ---- stage 1: suffix array by prefix doubling --------------------------------------------------
// The first sort key: the first letters of the rotation -- 13 letters of 3 bits in a u64, or, when the
// batch holds nothing but A,C,G,T, 12 letters of 2 bits in a u32 (three radix passes of 8 B per element
// instead of five of 12 B).  The set number is not in the key: the first sort is segmented by set.
struct InitKeyArgs { BatchView v; u64 *keys64; u32 *keys32; u32 *vals; int letters32; }; // letters32: 12 or 16 letters in a u32 key
HD void initkey_body(long long i, const InitKeyArgs &a) {
    u32 g = (u32)i;
    u32 k = seq_of(a.v, g);
    u32 off = LDG(a.v.seq_off + k), n = LDG(a.v.seq_off + k + 1) - off;
    u32 p = g - off;
    if (a.keys32) { // nothing but A,C,G,T in the batch: the 12 (16) letters are the top 24 (32) bits of one word of the packed text
        a.keys32[g] = (u32)(lexkey2(fetch2(a.v.p2, LDG(a.v.dbl_off + k) + p)) >> (64 - 2 * a.letters32));
    } else {
        u64 key = 0;
        for (int t = 0; t < CSA_K0; t++) {
            key = (key << CSA_LETTER_BITS) | a.v.code[off + p];
            if (++p == n) p = 0;
        }
        a.keys64[g] = key;
    }
    if (a.vals) a.vals[g] = g; // (the first pass of the first sort makes up the suffix numbers itself)
}
MAP_KERNEL_N(initkey, InitKeyArgs, 21)

// head[i] = i where a new group of equal keys starts, else 0 (max-scanned afterwards)
// With lcp != nullptr (first sort only) a border also gets its LCP: the letters the two keys share.
struct FlagArgs { const u64 *keys; const u32 *keys32; u32 *head; u32 *ngroups; u32 *lcp; int letters; int lbits;
                  const u32 *sa; const u32 *seqof; const u32 *seq_off; int clamp;
                  u32 base; }; // the arrays start at SA place `base` (a bucket of a sharded run), heads are SA places
HD bool flag_differs(const FlagArgs &a, long long i) {
    return a.keys32 ? a.keys32[i] != a.keys32[i - 1] : a.keys[i] != a.keys[i - 1];
}
HD void flag_lcp(const FlagArgs &a, long long i) { // keys differ (or i == 0): first letter sits in the top field
    u32 c = 0;
    if (i > 0) {
        const u64 d = a.keys32 ? (u64)(a.keys32[i] ^ a.keys32[i - 1]) : (a.keys[i] ^ a.keys[i - 1]);
        c = (u32)(CSA_CLZLL(d) - (64 - a.letters * a.lbits)) / (u32)a.lbits;
        if (a.clamp) { // a sequence shorter than the key: never more than the shorter rotation (gencycsuffixtrees.c:500)
            const u32 ka = a.seqof[a.sa[i - 1]], kb = a.seqof[a.sa[i]];
            const u32 na = a.seq_off[ka + 1] - a.seq_off[ka], nb = a.seq_off[kb + 1] - a.seq_off[kb];
            const u32 cap = na < nb ? na : nb;
            if (c > cap) c = cap;
        }
    }
    a.lcp[i] = c;
}
#ifdef CSA_EMU
HD void flag_body(long long i, const FlagArgs &a) {
    bool f = (i == 0) || flag_differs(a, i);
    a.head[i] = f ? a.base + (u32)i : 0u;
    if (f && a.lcp) flag_lcp(a, i);
    COUNT_IF(a.ngroups, f);
}
MAP_KERNEL(flag, FlagArgs, 12)
#else
// one atomic per CTA: same-address atomics serialise in L2 (one per warp cost 0.5 ms per launch)
__global__ void __launch_bounds__(256) k_flag(long long n, FlagArgs a) {
    long long i = (long long)blockIdx.x * (256 * MAP_ITEMS) + threadIdx.x;
    int mine = 0;
#pragma unroll 1
    for (int j = 0; j < MAP_ITEMS && i < n; j++, i += 256) {
        const bool f = (i == 0) || flag_differs(a, i);
        a.head[i] = f ? a.base + (u32)i : 0u;
        if (f && a.lcp) flag_lcp(a, i);
        mine += f ? 1 : 0;
    }
    __shared__ int s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_c, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_c) atomicAdd(a.ngroups, (u32)s_c);
}
static inline void launch_flag(Exec &ex, long long n, FlagArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_flag", 12.0 * n);
    k_flag<<<(unsigned)((n + 256 * MAP_ITEMS - 1) / (256 * MAP_ITEMS)), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// the first suffix of a set starts a group even when its key equals the last key of the set before
// (segmented first sort: the set number is not part of the key)
struct SetStartArgs { BatchView v; u32 *head; u32 *ngroups; u32 *lcp; };
HD void setstart_body(long long s, const SetStartArgs &a) {
    u32 pos = a.v.set_base0[s];
    if (pos != 0 && a.head[pos] != pos) { a.head[pos] = pos; ATOMIC_ADD(a.ngroups, 1u); }
    if (a.lcp) a.lcp[pos] = 0; // nothing in common with another set
}
MAP_KERNEL(setstart, SetStartArgs, 8)

struct SetRankArgs { const u32 *sa; const u32 *head; u32 *rank; };
HD void setrank_body(long long i, const SetRankArgs &a) { a.rank[a.sa[i]] = a.head[i]; }
MAP_KERNEL_N(setrank, SetRankArgs, 12)

// key of the doubling round: (rank of the first h letters, rank of the next h letters)
struct Key2Args { BatchView v; const u32 *sa; const u32 *rank; u64 *keys; u32 h; int nbits; };
HD void key2_body(long long i, const Key2Args &a) {
    u32 g = a.sa[i];
    u32 r1 = LDG(a.rank + g), r2 = LDG(a.rank + cyc_add(a.v, g, a.h));
    a.keys[i] = ((u64)r1 << a.nbits) | r2;
}
MAP_KERNEL(key2, Key2Args, 24)

// ---- stage 2: LCP of neighbouring suffixes, capped at the shorter rotation -------------------------
// gencycsuffixtrees.c:500: a path of the tree ends after textsize letters.

// Text order (Kasai): if rotation p shares L letters with its SA predecessor, rotation p+1 shares at
// least L-1 with its own, so a thread that walks LCP_CHUNK consecutive positions of a sequence
// extends one running match instead of starting every comparison at letter 0.  Cost per suffix: the
// inverse SA entry (coalesced), sa[r-1] and the predecessor's sequence (gathers), ~1-2 word compares
// on the L2-resident packed text, one scattered 4 B store.
#define LCP_CHUNK 32
struct IsaArgs { const u32 *sa; u32 *isa; };
HD void isa_body(long long i, const IsaArgs &a) { a.isa[a.sa[i]] = (u32)i; }
MAP_KERNEL_N(isa, IsaArgs, 12)

struct LcpArgs { BatchView v; const u32 *sa; const u32 *isa; u32 *lcp; const u32 *any_other; };
HD void lcp_body(long long t, const LcpArgs &a) {
    const u32 g0 = (u32)t * LCP_CHUNK;
    const u32 g1 = (g0 + LCP_CHUNK < a.v.N) ? g0 + LCP_CHUNK : a.v.N;
    const bool masks = *a.any_other != 0; // any letter outside ACGT in the batch?
    u32 k = CSA_NONE, off = 0, n = 0, s0 = 0;
    u64 dk = 0;
    u32 h = 0;
    // the chain inverse SA entry -> predecessor -> its sequence is three trips to memory per suffix: a three-stage
    // pipeline in registers -- the entry of the suffix three on, the predecessor of the one two on and the sequence
    // of the next one's predecessor are fetched while this suffix is compared
    u32 r_0 = a.isa[g0], r_1 = g0 + 1 < g1 ? a.isa[g0 + 1] : 0u, r_2 = g0 + 2 < g1 ? a.isa[g0 + 2] : 0u;
    u32 b_0 = r_0 > 0 ? a.sa[r_0 - 1] : 0u, b_1 = (g0 + 1 < g1 && r_1 > 0) ? a.sa[r_1 - 1] : 0u;
    u32 kb_0 = seq_of(a.v, b_0);
    for (u32 g = g0; g < g1; g++) {
        const u32 r_3 = g + 3 < g1 ? a.isa[g + 3] : 0u;
        const u32 b_2 = (g + 2 < g1 && r_2 > 0) ? a.sa[r_2 - 1] : 0u;
        const u32 kb_1 = g + 1 < g1 ? seq_of(a.v, b_1) : 0u;
        const u32 r = r_0, b = b_0, kb = kb_0;
        r_0 = r_1; r_1 = r_2; r_2 = r_3; b_0 = b_1; b_1 = b_2; kb_0 = kb_1;
        u32 kg = seq_of(a.v, g);
        if (kg != k) { // a new sequence begins: nothing carries over
            k = kg;
            off = LDG(a.v.seq_off + k);
            n = LDG(a.v.seq_off + k + 1) - off;
            s0 = LDG(a.v.set_base0 + LDG(a.v.seq_set + k));
            dk = LDG(a.v.dbl_off + k);
            h = 0;
        }
        if (r == s0) { a.lcp[r] = 0; h = 0; continue; } // first suffix of its set
        u32 ob = LDG(a.v.seq_off + kb), nb = LDG(a.v.seq_off + kb + 1) - ob;
        u32 cap = n < nb ? n : nb;
        if (h > cap) h = cap;
        u64 xa = dk + (g - off), xb = LDG(a.v.dbl_off + kb) + (b - ob);
        while (h < cap) {
            u64 d2 = fetch2(a.v.p2, xa + h) ^ fetch2(a.v.p2, xb + h);
            int f = 32;
            if (d2) f = ctz64(d2) >> 1;
            if (masks) {
                u32 dm = fetchm(a.v.pm, xa + h) ^ fetchm(a.v.pm, xb + h);
                if (dm) { int f2 = ctz32(dm); if (f2 < f) f = f2; }
            }
            if (f < 32) { h += (u32)f; break; }
            h += 32;
        }
        if (h > cap) h = cap;
        a.lcp[r] = h;
        if (h > 0) h--;
    }
}
MAP_KERNEL(lcp, LcpArgs, 20 * LCP_CHUNK)

// SA order: every pair compared from letter 0 -- coalesced reads and writes, no inverse array; the
// better choice when matches are short (a word compare or two per pair).  The host samples the
// LCP of a few thousand pairs with this same kernel (stride > 1, sum only) and picks.
struct LcpDirectArgs { BatchView v; const u32 *sa; u32 *lcp; const u32 *any_other; u32 stride; unsigned long long *sum; };
HD void lcpdirect_body(long long t, const LcpDirectArgs &a) {
    const u64 i64 = (u64)t * a.stride;
    if (i64 >= a.v.N) return;
    const u32 i = (u32)i64;
    u32 h = 0;
    if (i > 0) {
        u32 ga = a.sa[i - 1], gb = a.sa[i];
        u32 ka = seq_of(a.v, ga), kb = seq_of(a.v, gb);
        if (LDG(a.v.seq_set + ka) == LDG(a.v.seq_set + kb)) {
            const bool masks = *a.any_other != 0;
            u32 oa = LDG(a.v.seq_off + ka), ob = LDG(a.v.seq_off + kb);
            u32 na = LDG(a.v.seq_off + ka + 1) - oa, nb = LDG(a.v.seq_off + kb + 1) - ob;
            u32 cap = na < nb ? na : nb;
            u64 xa = LDG(a.v.dbl_off + ka) + (ga - oa), xb = LDG(a.v.dbl_off + kb) + (gb - ob);
            while (h < cap) {
                u64 d2 = fetch2(a.v.p2, xa + h) ^ fetch2(a.v.p2, xb + h);
                int f = 32;
                if (d2) f = ctz64(d2) >> 1;
                if (masks) {
                    u32 dm = fetchm(a.v.pm, xa + h) ^ fetchm(a.v.pm, xb + h);
                    if (dm) { int f2 = ctz32(dm); if (f2 < f) f = f2; }
                }
                if (f < 32) { h += (u32)f; break; }
                h += 32;
            }
            if (h > cap) h = cap;
        }
    }
    if (a.sum) { if (h > 4096) h = 4096; ATOMIC_ADD(a.sum, (unsigned long long)h); }
    else a.lcp[i] = h;
}
MAP_KERNEL(lcpdirect, LcpDirectArgs, 12)

// ---- stage 3: common blocks ------------------------------------------------------------------------
// R[l] = smallest r such that SA[l..r] holds a suffix of every sequence of the set (>= end of the
// set when there is none).  An LCP interval [lb,rb] "belongs to all the sequences"
// (node->fromseqs == allseqsmask, gencycsuffixtrees.c:34) iff rb >= R[lb].
//   R[l] = max( last first-occurrence of any colour, max_{j<l} next index of j's colour )
// The colour lists come from one stable radix pass of the SA indices by sequence-in-set.
struct ColorKeyArgs { BatchView v; const u32 *sa; u64 *keys; u32 *vals; };
HD void colorkey_body(long long i, const ColorKeyArgs &a) {
    u32 k = seq_of_few(a.v, a.sa[i]);
    // low bits (the only ones sorted on): sequence-in-set; high word: the sequence, carried along so
    // that k_next reads its neighbours' sequences from the sorted keys instead of gathering them
    a.keys[i] = ((u64)k << 32) | (k - LDG(a.v.set_seq0 + LDG(a.v.seq_set + k)));
    a.vals[i] = (u32)i;
}
MAP_KERNEL(colorkey, ColorKeyArgs, 20)

// after the sort: vals = SA indices ordered by (colour, index).  cover[i] (read at i+1) = next
// index of the same sequence; cover[set start] collects the latest first occurrence.
struct NextArgs { BatchView v; const u64 *keys; const u32 *vals; u32 *nxt; u32 *firstmax; };
HD void next_body(long long j, const NextArgs &a) {
    u32 i = a.vals[j];
    u32 k = (u32)(a.keys[j] >> 32);
    u32 s = LDG(a.v.seq_set + k);
    u32 nx = LDG(a.v.set_base0 + s + 1); // "none": the end of the set
    if ((u32)j + 1 < a.v.N && (u32)(a.keys[j + 1] >> 32) == k) nx = a.vals[j + 1];
    a.nxt[i] = nx;
    bool first = (j == 0) || (u32)(a.keys[j - 1] >> 32) != k;
    if (first) ATOMIC_MAX(a.firstmax + s, i);
}
MAP_KERNEL(next, NextArgs, 16)

struct CoverArgs { BatchView v; const u32 *sa; const u32 *nxt; const u32 *firstmax; u32 *cover; };
HD void cover_body(long long i, const CoverArgs &a) { a.cover[i] = i > 0 ? a.nxt[i - 1] : 0u; }
MAP_KERNEL(cover, CoverArgs, 8)
// ... and the first place of every set starts from the last first-occurrence of any colour (launched after k_cover,
// one thread per set: no search for the set of every place)
HD void coverstart_body(long long s, const CoverArgs &a) { a.cover[LDG(a.v.set_base0 + s)] = a.firstmax[s]; }
MAP_KERNEL(coverstart, CoverArgs, 8)

// blocks: LCP intervals of exactly m suffixes, one of every sequence, that cannot be extended to
// the left by one and the same letter (csamsa.c:64,80,283).  One thread per left border.
struct BlockFindArgs { BatchView v; const u32 *sa; const u32 *lcp; const u32 *R; u32 *isblock; u32 *depth; u32 mmax; };
HD unsigned letter_before_suffix(const BatchView &v, u32 g) {
    u32 k = seq_of_few(v, g);
    u32 off = LDG(v.seq_off + k), n = LDG(v.seq_off + k + 1) - off;
    u32 p = g - off;
    return v.code[off + (p == 0 ? n - 1 : p - 1)];
}
#ifdef CSA_EMU
HD void blockfind_body(long long i, const BlockFindArgs &a) {
    u32 lb = (u32)i;
    a.isblock[lb] = 0;
    u32 s = set_of_pos(a.v, lb);
    u32 s0 = LDG(a.v.set_base0 + s), s1 = LDG(a.v.set_base0 + s + 1);
    u32 m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
    if (lb + m > s1 || m < 2) return;
    u32 rb = lb + m - 1;
    if (a.R[lb] != rb) return; // m suffixes holding all m sequences: each exactly once
    long long outer_l = (lb == s0) ? -1 : (long long)a.lcp[lb];
    long long outer_r = (rb + 1 == s1) ? -1 : (long long)a.lcp[rb + 1];
    long long outer = outer_l > outer_r ? outer_l : outer_r;
    u32 inner = 0xFFFFFFFFu;
    for (u32 j = lb + 1; j <= rb; j++) { u32 l = a.lcp[j]; if (l < inner) inner = l; }
    if ((long long)inner <= outer) return; // not an LCP interval
    // removeSuffixNodes: every occurrence preceded by one and the same letter -> a longer block
    // holds it (csamsa.c:85 leaves a list holding only the root alone: depth 0)
    if (inner > 0) {
        bool same = true;
        unsigned c0 = 0;
        for (u32 j = lb; j <= rb && same; j++) {
            u32 g = a.sa[j];
            u32 k = seq_of(a.v, g);
            u32 off = LDG(a.v.seq_off + k), n = LDG(a.v.seq_off + k + 1) - off;
            u32 p = g - off;
            unsigned c = a.v.code[off + (p == 0 ? n - 1 : p - 1)];
            if (j == lb) c0 = c; else if (c != c0) same = false;
        }
        if (same) return;
    }
    a.isblock[lb] = 1;
    a.depth[lb] = inner;
}
MAP_KERNEL(blockfind, BlockFindArgs, 16)
#else
// The same test, warp-cooperative: every lane screens one left border with the cover array; the
// few candidates a warp finds are then examined by the whole warp, one suffix of the window per
// lane -- smallest inner lcp by a min-reduction, "all preceded by one letter" by a ballot -- so the
// m dependent gathers of a candidate run side by side instead of one after the other.
__global__ void __launch_bounds__(256) k_blockfind(long long n, BlockFindArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    u32 lb = (u32)i, s0 = 0, s1 = 0, m = 0;
    bool cand = false;
    if (i < n && a.R[lb] - lb < a.mmax) { // (a window wider than the largest set's m holds a repeat: no search for the set)
        u32 s = set_of_pos(a.v, lb);
        s0 = LDG(a.v.set_base0 + s); s1 = LDG(a.v.set_base0 + s + 1);
        m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
        cand = (lb + m <= s1 && m >= 2 && a.R[lb] == lb + m - 1); // m suffixes, every sequence once
    }
    u32 res = 0, resdepth = 0;
    unsigned todo = __ballot_sync(0xffffffffu, cand);
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1;
        const u32 clb = __shfl_sync(0xffffffffu, lb, src), cm = __shfl_sync(0xffffffffu, m, src);
        const u32 cs0 = __shfl_sync(0xffffffffu, s0, src), cs1 = __shfl_sync(0xffffffffu, s1, src);
        const u32 crb = clb + cm - 1;
        u32 inner = 0xFFFFFFFFu, c0 = 0xFFu;
        bool same = true;
        for (u32 j = clb + lane; j <= crb; j += 32) {
            if (j > clb) { u32 l = a.lcp[j]; if (l < inner) inner = l; }
            unsigned c = letter_before_suffix(a.v, a.sa[j]);
            if (c0 == 0xFFu) c0 = c; else if (c != c0) same = false;
        }
        inner = __reduce_min_sync(0xffffffffu, inner);
        // all letters alike: every lane's own letters alike, and all lanes that saw any agree
        const unsigned have = __ballot_sync(0xffffffffu, c0 != 0xFFu);
        const u32 first = __shfl_sync(0xffffffffu, c0, __ffs((int)have) - 1);
        const bool all_same = __all_sync(0xffffffffu, same && (c0 == 0xFFu || c0 == first));
        const long long outer_l = (clb == cs0) ? -1 : (long long)a.lcp[clb];
        const long long outer_r = (crb + 1 == cs1) ? -1 : (long long)a.lcp[crb + 1];
        const long long outer = outer_l > outer_r ? outer_l : outer_r;
        // an LCP interval (inner > outer) that no single letter extends to the left (csamsa.c:80;
        // :85 leaves a list holding only the root alone: depth 0)
        const bool ok = (long long)inner > outer && !(inner > 0 && all_same);
        if ((int)lane == src) { res = ok ? 1u : 0u; resdepth = inner; }
    }
    if (i < n) {
        a.isblock[lb] = res;
        if (res) a.depth[lb] = resdepth;
    }
}
static inline void launch_blockfind(Exec &ex, long long n, BlockFindArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_blockfind", 16.0 * n);
    k_blockfind<<<(unsigned)((n + 255) / 256), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// ---- the same blocks without the cover array (sets of up to BF2_MAXM = 256 sequences, no counts asked for) ------------------
// A block is a window of exactly m suffix-array places: an LCP interval (the smallest lcp inside above both lcps at its
// borders), one rotation of every sequence (m places, m different sequences), not preceded by one and the same letter
// everywhere.  The LCP conditions need the LCP array alone and leave few candidates; only those have their m sequences
// looked up.  This path skips k_colorkey, the colour sort, k_next, k_cover and the scan that builds R[] (3.4 ms of a
// 15.8 ms step); R[] is still built when the counts of csamsa.c:332,338 are asked for (k_windepth, k_plateau work on it)
// or a set holds more than 256 sequences.
#define BF2_MAXM 256u // sequences a set may hold on this path
struct BlockFind2Args { BatchView v; const u32 *sa; const u32 *lcp; u32 *isblock; u32 *depth; u32 mmax; u32 off; }; // off: first place of the launch (a rank's own range)
HD bool blockfind2_screen(const BlockFind2Args &a, u32 lb, u32 *s0_out, u32 *s1_out, u32 *m_out) {
    const u32 s = set_of_pos(a.v, lb);
    const u32 s0 = LDG(a.v.set_base0 + s), s1 = LDG(a.v.set_base0 + s + 1);
    const u32 m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
    *s0_out = s0; *s1_out = s1; *m_out = m;
    if (m < 2 || lb + m > s1) return false;
    const u32 rb = lb + m - 1;
    // the borders first: lcp rises behind the left one and falls behind the right one
    if (lb != s0 && a.lcp[lb] >= a.lcp[lb + 1]) return false;
    if (rb + 1 != s1 && a.lcp[rb + 1] >= a.lcp[rb]) return false;
    return true;
}
#ifdef CSA_EMU
HD void blockfind2_body(long long i, const BlockFind2Args &a) {
    const u32 lb = (u32)i + a.off;
    a.isblock[lb] = 0;
    u32 s0, s1, m;
    if (!blockfind2_screen(a, lb, &s0, &s1, &m)) return;
    const u32 rb = lb + m - 1;
    const long long outer_l = (lb == s0) ? -1 : (long long)a.lcp[lb];
    const long long outer_r = (rb + 1 == s1) ? -1 : (long long)a.lcp[rb + 1];
    const long long outer = outer_l > outer_r ? outer_l : outer_r;
    u32 inner = 0xFFFFFFFFu;
    for (u32 j = lb + 1; j <= rb; j++) { const u32 l = a.lcp[j]; if (l < inner) inner = l; }
    if ((long long)inner <= outer) return;
    u64 seen[BF2_MAXM / 64] = {0, 0, 0, 0};
    bool same = true, twice = false;
    unsigned c0 = 0;
    const u32 q0 = LDG(a.v.set_seq0 + set_of_pos(a.v, lb));
    for (u32 j = lb; j <= rb; j++) {
        const u32 g = a.sa[j], k = seq_of(a.v, g);
        if (seen[(k - q0) >> 6] >> ((k - q0) & 63u) & 1ull) twice = true;
        seen[(k - q0) >> 6] |= 1ull << ((k - q0) & 63u);
        const unsigned c = letter_before_suffix(a.v, g);
        if (j == lb) c0 = c; else if (c != c0) same = false;
    }
    if (twice) return; // a sequence twice, another one missing
    if (inner > 0 && same) return;                              // csamsa.c:80 (csamsa.c:85: depth 0 is left alone)
    a.isblock[lb] = 1;
    a.depth[lb] = inner;
}
MAP_KERNEL(blockfind2, BlockFind2Args, 12)
#else
// set of SA place i for the threads of one CTA: one search for the CTA's first place, then a step or two forward
__device__ __forceinline__ u32 set_of_pos_cta(const BatchView &v, long long i, u32 off, u32 *s_first) {
    if (threadIdx.x == 0) *s_first = set_of_pos(v, (u32)((long long)blockIdx.x * blockDim.x) + off);
    __syncthreads();
    u32 s = *s_first;
    while (s + 1 < (u32)v.nsets && (u32)i >= LDG(v.set_base0 + s + 1)) s++;
    return s;
}
#define BF2_COOP_M 16u // sets of more sequences than this: a lane reads the first BF2_SCAN LCPs of its window, the warp the rest
#define BF2_SCAN 8u
#define BF2_THREADS 256 // (64-thread CTAs measured: 1.32 against 1.26 ms -- no tail to cut here)
template <bool COOP> // COOP: the batch holds sets of more than BF2_COOP_M sequences (two kernels: the other one keeps its 32 registers)
__global__ void __launch_bounds__(BF2_THREADS, 8) k_blockfind2(long long n, BlockFind2Args a) {
    __shared__ u32 s_first;
    __shared__ u32 s_seen[BF2_THREADS / 32][BF2_MAXM / 32]; // (sets of more than 64 sequences: the sequences seen, a bit each)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const u32 lb = (u32)i + a.off;
    const u32 s = set_of_pos_cta(a.v, (i < n ? i : n - 1) + a.off, a.off, &s_first);
    u32 m = 0, inner = 0, outer1 = 0; // outer1: the larger lcp at the window's two borders + 1 (0: the set's range ends on both sides)
    bool cand = false;
    if (i < n) {
        // every lane screens its own window with the LCP array alone: borders first, then the smallest lcp inside (of a set of
        // more than BF2_COOP_M sequences: the first BF2_SCAN of them -- the warp reads the rest together, below)
        const u32 s0 = LDG(a.v.set_base0 + s), s1 = LDG(a.v.set_base0 + s + 1);
        m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
        if (m >= 2 && lb + m <= s1) {
            const u32 rb = lb + m - 1;
            // (the four LCPs at the window's two borders asked for together, not one after the other's test)
            const u32 la = a.lcp[lb], lb1 = a.lcp[lb + 1], lr = a.lcp[rb], lr1 = (rb + 1 == s1) ? 0u : a.lcp[rb + 1];
            const long long outer_l = (lb == s0) ? -1 : (long long)la;
            const long long outer_r = (rb + 1 == s1) ? -1 : (long long)lr1;
            if (outer_l < (long long)lb1 && outer_r < (long long)lr) {
                const long long outer = outer_l > outer_r ? outer_l : outer_r;
                outer1 = (u32)(outer + 1);
                inner = 0xFFFFFFFFu;
                const u32 je = (COOP && m > BF2_COOP_M) ? lb + BF2_SCAN : rb;
                for (u32 j = lb + 1; j <= je && (long long)inner > outer; j++) { const u32 l = a.lcp[j]; inner = l < inner ? l : inner; }
                cand = (long long)inner > outer;
            }
        }
    }
    u32 res = 0;
    unsigned todo = __ballot_sync(0xffffffffu, cand);
    while (todo) { // the LCP intervals of exactly m places: the whole warp looks up the m sequences and the letters before
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1;
        const u32 clb = __shfl_sync(0xffffffffu, lb, src), cm = __shfl_sync(0xffffffffu, m, src);
        u32 cinner = __shfl_sync(0xffffffffu, inner, src);
        const u32 cs = __shfl_sync(0xffffffffu, s, src);
        const u32 crb = clb + cm - 1, q0 = LDG(a.v.set_seq0 + cs);
        if (COOP && cm > BF2_COOP_M) { // the rest of the window's LCPs, a lane each
            const u32 couter1 = __shfl_sync(0xffffffffu, outer1, src);
            u32 l = 0xFFFFFFFFu;
            for (u32 j = clb + 1u + BF2_SCAN + lane; j <= crb; j += 32) { const u32 x = a.lcp[j]; l = x < l ? x : l; }
            l = __reduce_min_sync(0xffffffffu, l);
            cinner = l < cinner ? l : cinner;
            if ((int)lane == src) inner = cinner;
            if ((u64)cinner + 1u <= (u64)couter1) continue; // no LCP interval after all
        }
        u32 seen_lo = 0, seen_hi = 0, c0 = 0xFFu;
        bool same = true, twice = false;
        const bool wide = cm > 64u;
        u32 *seen = s_seen[threadIdx.x >> 5];
        if (wide) { if (lane < BF2_MAXM / 32u) seen[lane] = 0u; __syncwarp(); }
        // (the sequences of a set are about as long as one another: where a suffix's sequence should stand if they were
        // equally long, then a step or two -- one round of loads where a search takes log2(m) dependent ones)
        const u32 b0 = LDG(a.v.seq_off + q0);
        const float per = (float)cm / (float)(LDG(a.v.seq_off + q0 + cm) - b0);
        for (u32 j = clb + lane; j <= crb; j += 32) {
            // the suffix's sequence from the set's own few sequence starts, the letter before it from the packed
            // text: both sit in L1/L2, where the gathers seqof[g] and code[g-1] go to HBM for every suffix of a batch
            const u32 g = a.sa[j];
            u32 lo = q0 + (u32)((float)(g - b0) * per);
            lo = lo < q0 + cm ? lo : q0 + cm - 1u;
            while (LDG(a.v.seq_off + lo) > g) lo--;      // (g is a suffix of this set: seq_off[q0] <= g < seq_off[q0 + cm])
            while (LDG(a.v.seq_off + lo + 1) <= g) lo++;
            const u32 col = lo - q0, off = LDG(a.v.seq_off + lo), nk = LDG(a.v.seq_off + lo + 1) - off;
            if (wide) twice |= (atomicOr(seen + (col >> 5), 1u << (col & 31u)) >> (col & 31u) & 1u) != 0u;
            else if (col < 32u) seen_lo |= 1u << col; else seen_hi |= 1u << (col - 32u);
            const u64 xp = LDG(a.v.dbl_off + lo) + (g - off) + nk - 1u; // (the doubled text holds s s s[0..64): the letter before place p is at p+n-1)
            const unsigned c = (LDG(a.v.pm + (xp >> 5)) >> (xp & 31u) & 1u) ? 4u : (unsigned)(LDG(a.v.p2 + (xp >> 5)) >> (2u * (xp & 31u))) & 3u;
            if (c0 == 0xFFu) c0 = c; else if (c != c0) same = false;
        }
        seen_lo = __reduce_or_sync(0xffffffffu, seen_lo);
        seen_hi = __reduce_or_sync(0xffffffffu, seen_hi);
        const unsigned have = __ballot_sync(0xffffffffu, c0 != 0xFFu);
        const u32 first = __shfl_sync(0xffffffffu, c0, __ffs((int)have) - 1);
        const bool all_same = __all_sync(0xffffffffu, same && (c0 == 0xFFu || c0 == first));
        const bool distinct = wide ? !__any_sync(0xffffffffu, twice) : (u32)(__popc(seen_lo) + __popc(seen_hi)) == cm; // (m places: no sequence twice = every sequence once)
        if (wide) __syncwarp();
        const bool ok = distinct && !(cinner > 0 && all_same);
        if ((int)lane == src) res = ok ? 1u : 0u;
    }
    if (i < n) {
        a.isblock[lb] = res;
        if (res) a.depth[lb] = inner;
    }
}
static inline void launch_blockfind2(Exec &ex, long long n, BlockFind2Args a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_blockfind2", 12.0 * n);
    if (a.mmax > BF2_COOP_M) k_blockfind2<true><<<(unsigned)((n + BF2_THREADS - 1) / BF2_THREADS), BF2_THREADS, 0, ex.stream>>>(n, a);
    else k_blockfind2<false><<<(unsigned)((n + BF2_THREADS - 1) / BF2_THREADS), BF2_THREADS, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// compaction of the block borders; blocks come out in SA order, i.e. grouped by set
struct BlockEmitArgs {
    BatchView v; const u32 *sa; const u32 *isblock; const u32 *bidx; const u32 *depth;
    u32 *blk_lb; u32 *blk_depth; u32 *blk_set; u32 *set_nblocks;
};
HD void blockemit_body(long long i, const BlockEmitArgs &a) {
    if (!a.isblock[i]) return;
    u32 b = a.bidx[i];
    u32 s = set_of_pos(a.v, (u32)i);
    a.blk_lb[b] = (u32)i;
    a.blk_depth[b] = a.depth[i];
    a.blk_set[b] = s;
}
MAP_KERNEL_N(blockemit, BlockEmitArgs, 8)
// blocks per set, from the running count (one thread per set): the one number the host waits for before the emit
struct SetCountArgs { BatchView v; const u32 *isblock; const u32 *bidx; u32 *set_nblocks; };
HD void setcount_body(long long s, const SetCountArgs &a) {
    const u32 s0 = LDG(a.v.set_base0 + s), s1 = LDG(a.v.set_base0 + s + 1);
    a.set_nblocks[s] = s1 > s0 ? a.bidx[s1 - 1] + a.isblock[s1 - 1] - a.bidx[s0] : 0u;
}
MAP_KERNEL(setcount, SetCountArgs, 12)

// ---- one set sharded over the ranks of a job: the block candidates of a rank's own range ------------------------------------
// (csa_gpu_shard_blocks_begin / _finish)  A block is exchanged as a record: its place, its depth and the m suffixes of its
// window -- all the later stages read of the suffix array.
struct LcpAtArgs { LcpDirectArgs d; u32 pos; };
HD void lcpat_body(long long, const LcpAtArgs &a) { lcpdirect_body((long long)a.pos, a.d); }
MAP_KERNEL(lcpat, LcpAtArgs, 0)

struct RangeMinArgs { const u32 *x; u32 lo; u32 *out; }; // *out = min(*out, x[lo .. lo+n))
HD void rangemin_body(long long i, const RangeMinArgs &a) { ATOMIC_MIN(a.out, a.x[a.lo + (u32)i]); }
MAP_KERNEL(rangemin, RangeMinArgs, 4)

struct BlkRecArgs { const u32 *sa; const u32 *isblock; const u32 *bidx; const u32 *depth; u32 off; u32 m; u32 *rec; };
HD void blkrecpack_body(long long i0, const BlkRecArgs &a) {
    const u32 i = (u32)i0 + a.off;
    if (!a.isblock[i]) return;
    u32 *r = a.rec + (size_t)a.bidx[i] * (2u + a.m);
    r[0] = i; r[1] = a.depth[i];
    for (u32 j = 0; j < a.m; j++) r[2 + j] = a.sa[i + j];
}
MAP_KERNEL(blkrecpack, BlkRecArgs, 12)
struct BlkUnpackArgs { const u32 *rec; u32 m; u32 *sa; u32 *blk_lb; u32 *blk_depth; u32 *blk_set; };
HD void blkrecunpack_body(long long b, const BlkUnpackArgs &a) {
    const u32 *r = a.rec + (size_t)b * (2u + a.m);
    a.blk_lb[b] = r[0]; a.blk_depth[b] = r[1]; a.blk_set[b] = 0;
    for (u32 j = 0; j < a.m; j++) a.sa[r[0] + j] = r[2 + j];
}
MAP_KERNEL(blkrecunpack, BlkUnpackArgs, 12)

// ---- stage 4: order of the block list ----------------------------------------------------------------
// insertSortedItem (nodeslinkedlists.c:36) keeps the list by depth, descending, and puts a block
// BEFORE the blocks of equal depth met earlier; blocks are met in the DFS order of the tree,
// whose children are in creation order (addBranch, gencycsuffixtrees.c:193) == the order of the
// first occurrence, in sequence 0 of the set, of the child's string.  The DFS number of every
// rotation of sequence 0 is computed on the LCP-interval tree of sequence 0 alone:
//   dfs(leaf) = sum over the nodes v on the path leaf..root of before(v),
//   before(v) = leaves under the siblings of v whose first occurrence precedes v's.
struct Seq0FlagArgs { BatchView v; const u32 *sa; u32 *flag; u32 off; };
HD void seq0flag_body(long long i, const Seq0FlagArgs &a) { // (a rotation of sequence 0 of its set: a range check, no gather)
    const u32 s = set_of_pos(a.v, (u32)i + a.off), k0 = LDG(a.v.set_seq0 + s);
    const u32 g = a.sa[i + a.off];
    a.flag[i] = (g >= LDG(a.v.seq_off + k0) && g < LDG(a.v.seq_off + k0 + 1)) ? 1u : 0u;
}
MAP_KERNEL(seq0flag, Seq0FlagArgs, 12)

struct Seq0EmitArgs { BatchView v; const u32 *sa; const u32 *flag; const u32 *idx0; u32 *sa0; u32 *saidx0; u32 *leaf_set; u32 off; u32 *count; u32 n; };
HD void seq0emit_body(long long i, const Seq0EmitArgs &a) { // (flag, idx0: for the places [off, off+n) of the launch)
    if (a.count && (u32)i + 1 == a.n) *a.count = a.idx0[i] + a.flag[i];
    if (!a.flag[i]) return;
    const u32 t = a.idx0[i], s = set_of_pos(a.v, (u32)i + a.off);
    a.sa0[t] = a.sa[i + a.off] - LDG(a.v.seq_off + LDG(a.v.set_seq0 + s)); // (sequence 0 of the set: no look-up of the suffix's sequence)
    a.saidx0[t] = (u32)i + a.off;
    a.leaf_set[t] = s;
}
MAP_KERNEL(seq0emit, Seq0EmitArgs, 8)
#ifndef CSA_EMU
// the three kernels above in ONE pass over the suffix array (flag, running count by decoupled look-back as in k_scan_chain,
// write-out): 4 B read per place instead of 28 B moved
struct Seq0CompactArgs { BatchView v; const u32 *sa; u32 *sa0; u32 *saidx0; u32 *leaf_set; unsigned long long *state; u32 off; u32 *count; }; // places [off, off+n); count: how many came out
__global__ void __launch_bounds__(CS_THREADS) k_seq0compact(long long n, Seq0CompactArgs a) {
    __shared__ u32 sm[33];
    __shared__ u32 s_tile, s_prefix, s_set;
    if (threadIdx.x == 0) s_tile = (u32)atomicAdd(a.state, 1ull); // state[0]: next tile number; state[1+t]: tile t
    __syncthreads();
    const u32 tile = s_tile;
    volatile unsigned long long *st = a.state + 1;
    const long long tile0 = (long long)tile * CS_TILE, base = tile0 + (long long)threadIdx.x * CS_ITEMS;
    if (threadIdx.x == 0) s_set = set_of_pos(a.v, (u32)tile0 + a.off);
    __syncthreads();
    u32 s = s_set;
    u32 g[CS_ITEMS], sset[CS_ITEMS];
    u32 cnt = 0, flags = 0;
    // (the set's tables are read again only where a thread's places cross into the next set, not for every place)
    u32 nextb = s + 1 < (u32)a.v.nsets ? LDG(a.v.set_base0 + s + 1) : 0xFFFFFFFFu;
    u32 k0 = LDG(a.v.set_seq0 + s), off = LDG(a.v.seq_off + k0), offe = LDG(a.v.seq_off + k0 + 1);
#pragma unroll
    for (int j = 0; j < CS_ITEMS; j++) {
        const long long i = base + j;
        g[j] = 0; sset[j] = 0;
        if (i < n) {
            if ((u32)i + a.off >= nextb) {
                while (s + 1 < (u32)a.v.nsets && (u32)i + a.off >= LDG(a.v.set_base0 + s + 1)) s++;
                nextb = s + 1 < (u32)a.v.nsets ? LDG(a.v.set_base0 + s + 1) : 0xFFFFFFFFu;
                k0 = LDG(a.v.set_seq0 + s); off = LDG(a.v.seq_off + k0); offe = LDG(a.v.seq_off + k0 + 1);
            }
            const u32 x = a.sa[i + a.off];
            if (x >= off && x < offe) { flags |= 1u << j; cnt++; g[j] = x - off; sset[j] = s; }
        }
    }
    u32 total;
    const u32 excl = block_scan_excl(cnt, total, ScanSum(), sm);
    if (threadIdx.x == 0) {
        st[tile] = ((unsigned long long)(tile == 0 ? 2u : 1u) << 32) | total;
        if (tile == 0) s_prefix = 0;
    }
    if (tile > 0 && threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        u32 run = 0;
        long long look = (long long)tile - 1;
        for (;;) {
            const long long t = look - lane;
            unsigned long long w = (t >= 0) ? st[t] : (2ull << 32);
            while (__any_sync(0xffffffffu, (w >> 32) == 0)) w = (t >= 0) ? st[t] : (2ull << 32);
            const unsigned incl = __ballot_sync(0xffffffffu, (w >> 32) == 2);
            const int stop = incl ? (__ffs((int)incl) - 1) : 32;
            u32 part = ((int)lane <= stop) ? (u32)w : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            run += part;
            if (incl) break;
            look -= 32;
        }
        if (lane == 0) { s_prefix = run; st[tile] = (2ull << 32) | (run + total); }
    }
    __syncthreads();
    u32 t = s_prefix + excl;
    if (a.count && threadIdx.x == 0 && tile0 + CS_TILE >= n) *a.count = s_prefix + total; // (the last tile knows the sum)
#pragma unroll
    for (int j = 0; j < CS_ITEMS; j++)
        if (flags >> j & 1u) { a.sa0[t] = g[j]; a.saidx0[t] = (u32)(base + j) + a.off; a.leaf_set[t] = sset[j]; t++; }
}
#endif


// The same three arrays without a pass over the whole suffix array: the stable radix pass by colour of stage 3
// (k_colorkey) leaves the SA places of colour 0 -- sequence 0 of every set -- first, set by set and in SA order.
struct Seq0TakeArgs { BatchView v; const u32 *sa; const u32 *places; u32 *sa0; u32 *saidx0; u32 *leaf_set; };
HD void seq0take_body(long long t, const Seq0TakeArgs &a) {
    u32 i = a.places[t];
    u32 g = a.sa[i];
    u32 k = seq_of(a.v, g);
    a.sa0[t] = g - LDG(a.v.seq_off + k);
    a.saidx0[t] = i;
    a.leaf_set[t] = LDG(a.v.seq_set + k);
}
MAP_KERNEL(seq0take, Seq0TakeArgs, 24)

struct Lcp0Args { const u32 *lcp; const u32 *saidx0; const u32 *leaf_set; const u32 *z0; u32 *lcp0; };
HD void lcp0_body(long long t, const Lcp0Args &a) {
    u32 s = a.leaf_set[t];
    if ((u32)t == LDG(a.z0 + s)) { a.lcp0[t] = 0; return; }
    u32 mn = 0xFFFFFFFFu;
    for (u32 j = a.saidx0[t - 1] + 1; j <= a.saidx0[t]; j++) { u32 l = a.lcp[j]; if (l < mn) mn = l; }
    a.lcp0[t] = mn;
}
MAP_KERNEL(lcp0, Lcp0Args, 12)

// nearest smaller / smaller-or-equal values around every border t (z0 < t < z1) of the set.
// A pyramid of block minima (factor 32) over lcp0 turns every search into O(32 log_32 N0) steps
// instead of a walk as long as the node is wide (the root's children span a quarter of the set).
// lcp0 is 0 at the first leaf of every set: a sentinel that stops searches at the set's borders.
#define PYR_MAX 8
struct Pyramid { const u32 *lev[PYR_MAX]; u32 size[PYR_MAX]; int nlev; };
struct PyrArgs { const u32 *in; u32 *out; u32 nin; };
HD void pyr_body(long long i, const PyrArgs &a) {
    u32 b0 = (u32)i * 32, b1 = b0 + 32 < a.nin ? b0 + 32 : a.nin;
    u32 mn = 0xFFFFFFFFu;
    for (u32 j = b0; j < b1; j++) { u32 x = a.in[j]; if (x < mn) mn = x; }
    a.out[i] = mn;
}
MAP_KERNEL(pyr, PyrArgs, 4)

// largest index < i whose value is < v (strict) or <= v; CSA_NONE if there is none
HD u32 prev_below(const Pyramid &py, u32 i, u32 v, bool or_equal) {
    int level = 0;
    u32 pos = i;
    for (;;) {
        u32 b0 = pos & ~31u;
        u32 p = pos;
        while (p > b0) {
            u32 x = py.lev[level][p - 1];
            if (or_equal ? x <= v : x < v) {
                p--;
                while (level > 0) { // walk down to the rightmost qualifying entry
                    level--;
                    u32 c1 = p * 32 + 32 < py.size[level] ? p * 32 + 32 : py.size[level];
                    u32 c = c1;
                    for (;;) {
                        u32 y = py.lev[level][c - 1];
                        if (or_equal ? y <= v : y < v) break;
                        c--;
                    }
                    p = c - 1;
                }
                return p;
            }
            p--;
        }
        if (b0 == 0 || level + 1 >= py.nlev) return CSA_NONE;
        pos = b0 >> 5;
        level++;
    }
}
// smallest index > i whose value is < v; CSA_NONE if there is none
HD u32 next_below(const Pyramid &py, u32 i, u32 v) {
    int level = 0;
    u32 pos = i + 1; // first candidate
    for (;;) {
        u32 b1 = (pos + 31) & ~31u; // end of the block that holds pos (pos itself if it starts a block)
        if (b1 > py.size[level]) b1 = py.size[level];
        u32 p = pos;
        while (p < b1) {
            if (py.lev[level][p] < v) {
                while (level > 0) {
                    level--;
                    u32 c = p * 32;
                    while (!(py.lev[level][c] < v)) c++;
                    p = c;
                }
                return p;
            }
            p++;
        }
        if (level + 1 >= py.nlev) return CSA_NONE;
        pos = (pos + 31) >> 5; // first block to the right not yet looked at
        level++;
        if (pos >= py.size[level]) return CSA_NONE;
    }
}

// What the block order needs of the LCP-interval tree of sequence 0 -- which of two leaves the DFS of csamsa.c:64 meets
// first -- is answered from two pyramids of block minima (over lcp0 and over the positions sa0), without building the tree:
// the node where leaves a < b part has depth d = min lcp0(a+1..b]; the child of that node that holds a leaf t is the widest
// range of leaves round t inside which lcp0 stays above d; children are met in the order of their first occurrence in
// sequence 0 (creation order, gencycsuffixtrees.c:193) = the smallest sa0 of the range.  (Round 1 built the whole tree --
// nearest smaller values, parents, first occurrences by atomics, children sorted by a 5-pass radix sort, DFS numbers by
// pointer jumping over 2 N0 nodes, 3.4 ms of an 18 ms step -- to number ALL rotations of sequence 0, of which a few hundred
// per set, the blocks, were ever looked up.)
struct Seq0Q {
    u32 N0; const u32 *z0; const u32 *leaf_set; const u32 *saidx0;
    Pyramid lcp; // over lcp0
    Pyramid pos; // over sa0
};
// smallest value of level 0 in [lo, hi), lo < hi
HD u32 pyr_min(const Pyramid &py, u32 lo, u32 hi) {
    u32 m = 0xFFFFFFFFu;
    int level = 0;
    while (lo < hi) {
        while (lo < hi && (lo & 31u)) { const u32 x = py.lev[level][lo]; m = x < m ? x : m; lo++; }
        while (lo < hi && (hi & 31u)) { hi--; const u32 x = py.lev[level][hi]; m = x < m ? x : m; }
        if (lo >= hi) break;
        if (level + 1 >= py.nlev) { for (; lo < hi; lo++) { const u32 x = py.lev[level][lo]; m = x < m ? x : m; } break; }
        lo >>= 5; hi >>= 5; level++;
    }
    return m;
}
// first occurrence in sequence 0 of the string of the child, below a node of depth d, that holds leaf t
HD u32 seq0_child_minpos(const Seq0Q &q, u32 t, u32 d) {
    u32 pl = prev_below(q.lcp, t + 1, d, true); // (lcp0 is 0 at the first leaf of every set: found inside the set)
    if (pl == CSA_NONE) pl = 0;
    u32 pr = next_below(q.lcp, t, d + 1);
    if (pr == CSA_NONE) pr = q.N0;
    return pyr_min(q.pos, pl, pr);
}
// does the DFS of csamsa.c:64 meet leaf t1 before leaf t2 (two leaves of one set)?
HD bool seq0_before(const Seq0Q &q, u32 t1, u32 t2) {
    if (t1 == t2) return false;
    const u32 a = t1 < t2 ? t1 : t2, b = t1 < t2 ? t2 : t1;
    const u32 d = pyr_min(q.lcp, a + 1, b + 1);
    const bool a_first = seq0_child_minpos(q, a, d) < seq0_child_minpos(q, b, d);
    return t1 < t2 ? a_first : !a_first;
}
// the first leaf of set s at or behind SA place i
HD u32 seq0_leaf_at(const Seq0Q &q, u32 s, u32 i) {
    const u32 z = LDG(q.z0 + s), n0 = LDG(q.z0 + s + 1) - z;
    u32 lo = 0, hi = n0;
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (q.saidx0[z + mid] < i) lo = mid + 1; else hi = mid; }
    return z + lo;
}

// per block: its leaf in sequence 0 and the key (set, depth descending); after the (stable) sort by that key the blocks of
// equal set and depth are put in the order the reference's list has them: the one the DFS meets LATER first
// (nodeslinkedlists.c:36 inserts in front of equal depths)
struct BlockKeyArgs {
    BatchView v; const u32 *sa; Seq0Q q;
    const u32 *blk_lb; const u32 *blk_depth; const u32 *blk_set;
    u64 *keys; u32 *vals; u32 *blk_leaf;
};
HD void blockkey_body(long long b, const BlockKeyArgs &a) {
    const u32 lb = a.blk_lb[b], s = a.blk_set[b];
    const u32 k0 = LDG(a.v.set_seq0 + s), m = LDG(a.v.set_seq0 + s + 1) - k0;
    for (u32 j = lb; j < lb + m; j++)
        if (seq_of(a.v, a.sa[j]) == k0) a.blk_leaf[b] = seq0_leaf_at(a.q, s, j); // the block's place in sequence 0
    a.keys[b] = ((u64)s << 32) | (0xFFFFFFFFu - a.blk_depth[b]);
    a.vals[b] = (u32)b;
}
MAP_KERNEL(blockkey, BlockKeyArgs, 16)

// cstart[i] = first place of the class (equal set and depth) of sorted place i: flags here, a max-scan by the host
struct BlockClassArgs { const u64 *keys; u32 *cstart; };
HD void blockclass_body(long long i, const BlockClassArgs &a) { a.cstart[i] = (i > 0 && a.keys[i - 1] != a.keys[i]) ? (u32)i : 0u; }
MAP_KERNEL(blockclass, BlockClassArgs, 12)

// Two blocks of a class mostly part high up in the tree (two random places of a 5 Mb genome share ~11 letters): the first
// occurrences of a block's ancestors' children down to depth BR_DEPTHS are looked up once per block (k_blocktab), a
// comparison is then one range minimum (where the two leaves part) and two table reads.
#define BR_DEPTHS 16
struct BlockTabArgs { Seq0Q q; const u32 *blk_leaf; u32 *tab; };
HD void blocktab_body(long long x, const BlockTabArgs &a) {
    const u32 b = (u32)(x / BR_DEPTHS), d = (u32)(x % BR_DEPTHS);
    a.tab[x] = seq0_child_minpos(a.q, a.blk_leaf[b], d);
}
MAP_KERNEL(blocktab, BlockTabArgs, 4)

struct BlockRankArgs { Seq0Q q; const u64 *keys; const u32 *vals; const u32 *blk_leaf; u32 B; const u32 *cstart; const u32 *tab; u32 *order; };
// does the DFS meet block x (leaf tx) before block y?
HD bool block_before(const BlockRankArgs &a, u32 x, u32 tx, u32 y, u32 ty) {
    const u32 lo = tx < ty ? tx : ty, hi = tx < ty ? ty : tx;
    const u32 d = pyr_min(a.q.lcp, lo + 1, hi + 1);
    if (d < (u32)BR_DEPTHS) return a.tab[(size_t)x * BR_DEPTHS + d] < a.tab[(size_t)y * BR_DEPTHS + d];
    return seq0_child_minpos(a.q, tx, d) < seq0_child_minpos(a.q, ty, d);
}
#ifdef CSA_EMU
HD void blockrank_body(long long i, const BlockRankArgs &a) {
    const u64 key = a.keys[i];
    const u32 me = a.vals[i], leaf = a.blk_leaf[me], c0 = a.cstart[i];
    u32 later = 0;
    for (u32 j = c0; j < a.B && a.keys[j] == key; j++)
        if (j != (u32)i && block_before(a, me, leaf, a.vals[j], a.blk_leaf[a.vals[j]])) later++;
    a.order[c0 + later] = me;
}
MAP_KERNEL(blockrank, BlockRankArgs, 16)
#else
// one WARP per block: the lanes take the other blocks of its class in turn (a class can hold thousands of blocks -- the
// short ones of a bacterial set)
__global__ void __launch_bounds__(256) k_blockrank(long long n, BlockRankArgs a) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (i >= n) return;
    const u64 key = a.keys[i];
    const u32 me = a.vals[i], leaf = a.blk_leaf[me], c0 = a.cstart[i];
    u32 later = 0;
    for (u32 j0 = c0;; j0 += 32u) {
        const u32 j = j0 + lane;
        const bool in = j < a.B && a.keys[j] == key;
        if (in && j != (u32)i) { const u32 other = a.vals[j]; if (block_before(a, me, leaf, other, a.blk_leaf[other])) later++; }
        if (!__all_sync(0xffffffffu, in)) break;
    }
    later = __reduce_add_sync(0xffffffffu, later);
    if (lane == 0) a.order[c0 + later] = me;
}
static inline void launch_blockrank(Exec &ex, long long n, BlockRankArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_blockrank", 16.0 * n);
    k_blockrank<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// blocks in list order: gather fields, write positions
struct BlockGatherArgs {
    BatchView v; const u32 *sa; const u32 *order; const u32 *blk_lb; const u32 *blk_depth; const u32 *blk_set;
    const u32 *set_blk0; const u32 *set_pos0;
    u32 *o_depth; u32 *o_set; int *o_pos;
};
HD u32 pos_offset(const u32 *set_blk0, const u32 *set_pos0, u32 s, u32 m, u32 b) {
    return LDG(set_pos0 + s) + (b - LDG(set_blk0 + s)) * m;
}
HD void blockgather_body(long long b, const BlockGatherArgs &a) {
    u32 ob = a.order[b];
    u32 s = a.blk_set[ob];
    u32 q0 = LDG(a.v.set_seq0 + s);
    u32 m = LDG(a.v.set_seq0 + s + 1) - q0;
    a.o_depth[b] = a.blk_depth[ob];
    a.o_set[b] = s;
    u32 po = pos_offset(a.set_blk0, a.set_pos0, s, m, (u32)b);
    u32 lb = a.blk_lb[ob];
    for (u32 j = lb; j < lb + m; j++) {
        u32 g = a.sa[j];
        u32 k = seq_of(a.v, g);
        a.o_pos[po + (k - q0)] = (int)(g - LDG(a.v.seq_off + k));
    }
}
MAP_KERNEL(blockgather, BlockGatherArgs, 16)

// ---- stage 5: chaining (collectNodeChains, csamsa.c:135-279) -----------------------------------------
// csamsa.c:147-183 walks every sequence round the circle and notes which block follows which; a
// block is noticed where it ENDS (position+depth, unrolled).  Two blocks are linked when the same
// successor follows in every sequence that has one.  Sorting the block ends of every sequence
// gives the same successor relation.
struct EndKeyArgs {
    BatchView v; const u32 *o_depth; const u32 *o_set; const int *o_pos; const u32 *set_blk0; const u32 *set_pos0;
    const u32 *elem_blk; // [E] block of element x (E = sum over blocks of m)
    u64 *keys; u32 *vals; int ebits;
};
HD void endkey_body(long long x, const EndKeyArgs &a) {
    u32 b = a.elem_blk[x];
    u32 s = a.o_set[b];
    u32 q0 = LDG(a.v.set_seq0 + s);
    u32 m = LDG(a.v.set_seq0 + s + 1) - q0;
    u32 po = pos_offset(a.set_blk0, a.set_pos0, s, m, b);
    u32 k = (u32)x - po;
    u32 e = (u32)a.o_pos[x] + a.o_depth[b];
    a.keys[x] = ((u64)(q0 + k) << a.ebits) | e;
    a.vals[x] = b;
}
MAP_KERNEL(endkey, EndKeyArgs, 24)

struct ElemBlkArgs { BatchView v; const u32 *o_set; const u32 *set_blk0; const u32 *set_pos0; u32 *elem_blk; };
HD void elemblk_body(long long b, const ElemBlkArgs &a) {
    u32 s = a.o_set[b];
    u32 m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
    u32 po = pos_offset(a.set_blk0, a.set_pos0, s, m, (u32)b);
    for (u32 k = 0; k < m; k++) a.elem_blk[po + k] = (u32)b;
}
MAP_KERNEL(elemblk, ElemBlkArgs, 8)

struct SegHeadArgs { const u64 *keys; u32 *seghead; int ebits; };
HD void seghead_body(long long j, const SegHeadArgs &a) {
    bool f = (j == 0) || (a.keys[j] >> a.ebits) != (a.keys[j - 1] >> a.ebits);
    a.seghead[j] = f ? (u32)j : 0u;
}
MAP_KERNEL(seghead, SegHeadArgs, 12)

struct LinkArgs {
    BatchView v; const u64 *keys; const u32 *vals; const u32 *seghead; const u32 *o_depth; int ebits;
    u32 *succ_lo; u32 *succ_hi;
};
HD void link_body(long long j, const LinkArgs &a) {
    u32 f = a.seghead[j];
    if ((u32)j == f) return;
    u64 mask = (1ull << a.ebits) - 1;
    u32 k = (u32)(a.keys[j] >> a.ebits);
    u32 n = LDG(a.v.seq_off + k + 1) - LDG(a.v.seq_off + k);
    u64 e0 = a.keys[f] & mask;
    if (e0 >= n) return; // csamsa.c:160: no block noticed in the first turn
    u64 limit = (u64)n + (e0 - a.o_depth[a.vals[f]]); // csamsa.c:168
    u64 e = a.keys[j] & mask;
    if (e >= limit) return;
    u32 prev = a.vals[j - 1], cur = a.vals[j];
    ATOMIC_MIN(a.succ_lo + prev, cur);
    ATOMIC_MAX(a.succ_hi + prev, cur);
}
MAP_KERNEL(link, LinkArgs, 20)

// next block and the gap to it: the smallest gap over the sequences (csamsa.c:199-207)
struct GapArgs {
    BatchView v; const u32 *succ_lo; const u32 *succ_hi; const u32 *o_depth; const u32 *o_set; const int *o_pos;
    const u32 *set_blk0; const u32 *set_pos0; int max_interval; int *next; int *gap;
};
HD void gap_body(long long b, const GapArgs &a) {
    u32 lo = a.succ_lo[b], hi = a.succ_hi[b];
    a.gap[b] = 0;
    if (lo == CSA_NONE || lo != hi) { a.next[b] = -1; return; }
    u32 cur = lo;
    u32 s = a.o_set[b];
    u32 q0 = LDG(a.v.set_seq0 + s);
    u32 m = LDG(a.v.set_seq0 + s + 1) - q0;
    u32 pp = pos_offset(a.set_blk0, a.set_pos0, s, m, (u32)b), pc = pos_offset(a.set_blk0, a.set_pos0, s, m, cur);
    long long iv = LLONG_MAX;
    for (u32 k = 0; k < m; k++) {
        long long count = 0;
        int posc = a.o_pos[pc + k], posp = a.o_pos[pp + k];
        if (posc < posp) count += (long long)(LDG(a.v.seq_off + q0 + k + 1) - LDG(a.v.seq_off + q0 + k));
        count += (long long)posc - ((long long)posp + (long long)a.o_depth[b]);
        if (count < iv) iv = count;
    }
    if (iv > (long long)a.max_interval) { a.next[b] = -1; return; }
    a.next[b] = (int)cur;
    a.gap[b] = (int)iv;
}
MAP_KERNEL(gap, GapArgs, 16)

// csamsa.c:185-233, literally, one thread per set (the walk is a chain of dependent steps)
struct ChainArgs {
    const u32 *set_blk0; const u32 *o_depth; const int *next; const int *gap;
    int *size; int *total; int *interval; u32 *set_nchains; u32 *set_flags;
    int skip_big; // sets of CH_CAP < blocks <= CHB_MAX are k_chain_big's
};
// the walk over one set's blocks; arrays are indexed from the set's first block (b0), `next` holds
// batch-wide block numbers
HD void chain_walk(u32 b0, u32 B, const u32 *depth, const int *next, const int *gap, int *size, int *total,
                   int *interval, u32 *nchains, bool *hangs) {
    u32 mcs = B;
    long long guard_max = 4ll * B + 16;
    bool hang = false;
    for (u32 b = 0; b < B && !hang; b++) {
        if (total[b] == -1) continue;
        size[b] = (int)depth[b];
        u32 prev = b;
        int cur = next[b];
        long long guard = 0;
        while (cur != -1) {
            if (++guard > guard_max) { hang = true; break; }
            u32 c = (u32)cur - b0;
            int iv = gap[prev];
            if (total[c] > 0) {
                size[b] += size[c];
                total[b] += total[c];
                interval[prev] = iv;
                total[b] += iv;
                size[c] = (int)depth[c];
                total[c] = -1;
                mcs--;
                break;
            }
            size[c] = (int)depth[c];
            size[b] += size[c];
            interval[prev] = iv;
            total[b] += iv;
            total[c] = -1;
            mcs--;
            prev = c;
            cur = next[c];
        }
        total[b] += size[b];
    }
    *nchains = mcs;
    *hangs = hang;
}
#ifdef CSA_EMU
HD void chain_body(long long s, const ChainArgs &a) {
    u32 b0 = a.set_blk0[s], B = a.set_blk0[s + 1] - b0;
    u32 mcs;
    bool hang;
    chain_walk(b0, B, a.o_depth + b0, a.next + b0, a.gap + b0, a.size + b0, a.total + b0, a.interval + b0, &mcs, &hang);
    a.set_nchains[s] = mcs;
    if (hang) ATOMIC_OR(a.set_flags + s, 2u);
}
MAP_KERNEL(chain, ChainArgs, 16)
#else
// one CTA per set: the walk is a chain of dependent loads, so the set's block arrays are brought
// into shared memory first (a mitogenome set has a few hundred blocks) and one thread walks them
// there; sets with more blocks than fit are walked in global memory
#define CH_THREADS 128
#define CH_CAP 1536
#define CHB_MAX 65534u
__global__ void __launch_bounds__(CH_THREADS) k_chain(ChainArgs a) {
    __shared__ u32 s_depth[CH_CAP];
    __shared__ int s_next[CH_CAP], s_gap[CH_CAP], s_size[CH_CAP], s_total[CH_CAP], s_interval[CH_CAP];
    const u32 s = blockIdx.x;
    const u32 b0 = a.set_blk0[s], B = a.set_blk0[s + 1] - b0;
    u32 mcs = 0;
    bool hang = false;
    if (B <= CH_CAP) {
        for (u32 i = threadIdx.x; i < B; i += CH_THREADS) {
            s_depth[i] = a.o_depth[b0 + i]; s_next[i] = a.next[b0 + i]; s_gap[i] = a.gap[b0 + i];
            s_size[i] = 0; s_total[i] = 0; s_interval[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) chain_walk(b0, B, s_depth, s_next, s_gap, s_size, s_total, s_interval, &mcs, &hang);
        __syncthreads();
        for (u32 i = threadIdx.x; i < B; i += CH_THREADS) {
            a.size[b0 + i] = s_size[i]; a.total[b0 + i] = s_total[i]; a.interval[b0 + i] = s_interval[i];
        }
    } else if (a.skip_big && B <= CHB_MAX) {
        return;
    } else if (threadIdx.x == 0) {
        chain_walk(b0, B, a.o_depth + b0, a.next + b0, a.gap + b0, a.size + b0, a.total + b0, a.interval + b0, &mcs, &hang);
    }
    if (threadIdx.x == 0) {
        a.set_nchains[s] = mcs;
        if (hang) atomicOr(a.set_flags + s, 2u);
    }
}
static inline void launch_chain(Exec &ex, long long nsets, ChainArgs a) {
    if (nsets <= 0) return;
    PROF_BEGIN(ex, "k_chain", 0.0);
    k_chain<<<(unsigned)nsets, CH_THREADS, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
#endif

#ifndef CSA_EMU
// A set with tens of thousands of blocks (bacterial chromosomes): one thread walking them in global memory
// pays an L2 round trip per block (31 ms for 50 000 blocks).  Only the ORDER of the walk is sequential --
// which block a walk takes next and whether that block heads a finished chain -- and that needs nothing but
// `next` and one state byte per block, which fit shared memory (3 B per block).  So:
//   1. thread 0 replays the control flow of collectNodeChains (csamsa.c:185-233) in shared memory and logs
//      every step as an event (walk's head, block before, block, kind);
//   2. all threads add up gaps and depths per walk from the events (the sums of :203-226);
//   3. sizes/totals of walks that took over a finished chain follow from that chain's (a few parallel rounds);
//   4. the two places where the control flow looks at a SUM (total[c] > 0 at :201 for a finished head, and for
//      the walk's own head when a circular chain closes) were taken as true in 1. and are checked now; a set
//      that fails the check (never seen: every term depth + gap is positive) is redone by the literal walk.
#define CHB_THREADS 1024
#define CHB_ABS 1u    // total == -1: taken by some walk
#define CHB_HEAD 2u   // has finished its own walk
struct ChainBigArgs {
    ChainArgs c;
    const u32 *sets; // the sets this launch handles
    unsigned long long *events; u32 *ev_base; // per set: first event slot; capacity 4 * blocks
    int *wS, *wT, *wchild; u32 *wdone;        // per block scratch
    u32 *wpred, *wjump, *wmin, *wjump2, *wmin2; // per block scratch of the parallel walk order (jump/min double buffered)
    u32 *redo;                                // per set: 1 = the literal walk must redo it
};
__global__ void __launch_bounds__(CHB_THREADS) k_chain_big(ChainBigArgs a) {
    extern __shared__ unsigned char chb_dyn[];
    __shared__ u32 s_nev, s_mcs, s_fail, s_hang, s_multi;
    const u32 s = a.sets[blockIdx.x];
    const u32 b0 = a.c.set_blk0[s], B = a.c.set_blk0[s + 1] - b0;
    unsigned short *nx = (unsigned short *)chb_dyn;
    unsigned char *st = chb_dyn + 2 * (size_t)((B + 1) & ~1u);
    const u32 *depth = a.c.o_depth + b0;
    const int *gap = a.c.gap + b0;
    int *size = a.c.size + b0, *total = a.c.total + b0, *interval = a.c.interval + b0;
    int *wS = a.wS + b0, *wT = a.wT + b0, *wchild = a.wchild + b0;
    u32 *wdone = a.wdone + b0, *wpred = a.wpred + b0, *wjump = a.wjump + b0, *wmin = a.wmin + b0, *wjump2 = a.wjump2 + b0, *wmin2 = a.wmin2 + b0;
    unsigned long long *ev = a.events + a.ev_base[blockIdx.x];
    const u32 evcap = 4u * B;
    for (u32 i = threadIdx.x; i < B; i += CHB_THREADS) {
        const int n = a.c.next[b0 + i];
        nx[i] = n < 0 ? (unsigned short)0xFFFFu : (unsigned short)((u32)n - b0);
        st[i] = 0;
        wS[i] = 0; wT[i] = 0; wchild[i] = -1; wdone[i] = 0; // wS, wT: sums of depths / gaps of the walk that starts at i
        wpred[i] = CSA_NONE;
    }
    if (threadIdx.x == 0) { s_nev = 0; s_mcs = B; s_fail = 0; s_hang = 0; s_multi = 0; }
    __syncthreads();
    // ---- 1a. the usual shape: no block is the `next` of two others, the blocks form paths and rings ----
    // Then the walk order needs no walking.  On a path, a block starts a walk iff its number is smaller than
    // every number upstream of it (the outer loop of :185 reaches it before anything that could take it);
    // every other block, and every such head but the first, is taken by the walk of the smallest number
    // upstream.  On a ring the smallest number walks all the way round and closes on itself.  "Smallest
    // number upstream" is a prefix minimum along the paths: pointer jumping, log2(blocks) rounds, all threads.
    for (u32 i = threadIdx.x; i < B; i += CHB_THREADS)
        if (nx[i] != 0xFFFFu && atomicCAS(&wpred[nx[i]], CSA_NONE, i) != CSA_NONE) s_multi = 1;
    __syncthreads();
    if (!s_multi) {
        for (u32 i = threadIdx.x; i < B; i += CHB_THREADS) { wjump[i] = wpred[i]; wmin[i] = wpred[i]; }
        __syncthreads();
        u32 rounds = 1;
        while ((1u << rounds) < B) rounds++;
        for (u32 r = 0; r <= rounds; r++) { // a round reads the pointers of the round before only
            bool moved = false;
            for (u32 i = threadIdx.x; i < B; i += CHB_THREADS) {
                const u32 j = wjump[i];
                u32 mi = wmin[i], ji = CSA_NONE;
                if (j != CSA_NONE) {
                    const u32 mj = wmin[j];
                    mi = mj < mi ? mj : mi;
                    ji = wjump[j];
                    moved = true;
                }
                wmin2[i] = mi; wjump2[i] = ji;
            }
            u32 *t1 = wjump; wjump = wjump2; wjump2 = t1;
            u32 *t2 = wmin; wmin = wmin2; wmin2 = t2;
            if (!__syncthreads_or(moved)) break;
        }
        u32 nev = 0;
        for (u32 c = threadIdx.x; c < B; c += CHB_THREADS) {
            const u32 pr = wpred[c];
            if (pr == CSA_NONE) { st[c] = CHB_HEAD; continue; } // first block of a path
            const int iv = gap[pr];
            interval[pr] = iv;
            nev++;
            const bool ring = wjump[c] != CSA_NONE;
            const u32 m = wmin[c]; // smallest number upstream (on a ring: of the whole ring, c included)
            if (ring && m == c) { atomicAdd(&wT[c], iv); wchild[c] = -2 - (int)pr; st[c] = CHB_HEAD; }
            else if (!ring && c < m) { atomicAdd(&wT[m], iv); wchild[m] = (int)c; st[c] = CHB_HEAD | CHB_ABS; }
            else { atomicAdd(&wT[m], iv); atomicAdd(&wS[m], (int)depth[c]); st[c] = CHB_ABS; }
        }
        if (nev) atomicAdd(&s_nev, nev);
        __syncthreads();
        if (threadIdx.x == 0) s_mcs = B - s_nev;
        __syncthreads();
    }
    // ---- 1b. any other shape: thread 0 replays the order of the walks, then 2. adds up ----
    if (s_multi && threadIdx.x == 0) {
        u32 nev = 0, mcs = B;
        const long long guard_max = 4ll * B + 16;
        bool fail = false, hang = false;
        for (u32 b = 0; b < B && !hang && !fail; b++) {
            if (st[b] & CHB_ABS) continue;
            u32 prev = b, cur = nx[b];
            long long guard = 0;
            while (cur != 0xFFFFu) {
                if (++guard > guard_max) { hang = true; break; }
                if (nev == evcap) { fail = true; break; }
                const u32 c = cur;
                const bool finished = (c == b) || ((st[c] & CHB_HEAD) && !(st[c] & CHB_ABS)); // total[c] > 0 (checked in 4.)
                ev[nev++] = ((unsigned long long)b << 40) | ((unsigned long long)prev << 20) | ((unsigned long long)c << 2) | (finished ? (c == b ? 2ull : 1ull) : 0ull);
                mcs--;
                if (finished) {
                    if (c != b) { st[c] |= CHB_ABS; wchild[b] = (int)c; } else wchild[b] = -2 - (int)prev; // (closed on itself after `prev`)
                    break;
                }
                st[c] |= CHB_ABS;
                prev = c;
                cur = nx[c];
            }
            st[b] |= CHB_HEAD;
        }
        s_nev = nev; s_mcs = mcs; s_fail = fail ? 1u : 0u; s_hang = hang ? 1u : 0u;
    }
    __syncthreads();
    if (s_multi && (s_hang || s_fail)) {
        if (threadIdx.x == 0) {
            if (s_hang) atomicOr(a.c.set_flags + s, 2u); else a.redo[s] = 1;
            a.c.set_nchains[s] = s_mcs;
        }
        return;
    }
    // ---- 2. sums per walk ----
    const u32 nev = s_multi ? s_nev : 0u;
    for (u32 e = threadIdx.x; e < nev; e += CHB_THREADS) {
        const unsigned long long w = ev[e];
        const u32 b = (u32)(w >> 40), prev = (u32)(w >> 20) & 0xFFFFFu, c = (u32)(w >> 2) & 0x3FFFFu, kind = (u32)w & 3u;
        const int iv = gap[prev];
        interval[prev] = iv;
        atomicAdd(&wT[b], iv);
        if (kind == 0) atomicAdd(&wS[b], (int)depth[c]);
    }
    __syncthreads();
    // ---- 3. sizes and totals: S = own depths (+ S of the chain taken over), T = own gaps (+ T of that chain) + S ----
    // (a round only builds on chains finished in an EARLIER round: wdone 2 = worked out in this round, 1 = before)
    for (;;) {
        for (u32 b = threadIdx.x; b < B; b += CHB_THREADS) {
            if (!(st[b] & CHB_HEAD) || wdone[b]) continue;
            const int ch = wchild[b];
            if (ch >= 0 && wdone[ch] != 1u) continue; // the chain it took over is not worked out yet
            int S = (int)depth[b] + wS[b], T = wT[b];
            if (ch <= -2) { // circular chain (:201 with cur == b): what the reference leaves
                const int before = T - gap[-2 - ch]; // total[b] when the walk came back to b
                if (before <= 0) s_fail = 1;
                S = (int)depth[b]; T = (int)depth[b] - 1;
            } else {
                if (ch >= 0) { if (wT[ch] <= 0) s_fail = 1; S += wS[ch]; T += wT[ch]; }
                T += S;
            }
            wS[b] = S; wT[b] = T;
            wdone[b] = 2u;
        }
        __syncthreads();
        bool changed = false;
        for (u32 b = threadIdx.x; b < B; b += CHB_THREADS)
            if (wdone[b] == 2u) { wdone[b] = 1u; changed = true; }
        if (!__syncthreads_or(changed)) break;
    }
    if (s_fail) { if (threadIdx.x == 0) a.redo[s] = 1; return; }
    // ---- what the reference leaves in the list ----
    for (u32 i = threadIdx.x; i < B; i += CHB_THREADS) {
        if (st[i] & CHB_ABS) { size[i] = (int)depth[i]; total[i] = -1; }
        else { size[i] = wS[i]; total[i] = wT[i]; }
    }
    if (threadIdx.x == 0) a.c.set_nchains[s] = s_mcs;
}
static inline int launch_chain_big(Exec &ex, u32 nbig, size_t smem, ChainBigArgs a) {
    if (nbig == 0) return 0;
    // (per device and cheap: set on every launch rather than remembered per process)
    CUDA_TRY(cudaFuncSetAttribute(k_chain_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PROF_BEGIN(ex, "k_chain_big", 0.0);
    k_chain_big<<<nbig, CHB_THREADS, smem, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
    return 0;
}
#endif

// sortList (nodeslinkedlists.c:59): stable, by chain size, descending
struct SizeKeyArgs { const u32 *o_set; const int *size; u64 *keys; u32 *vals; };
HD void sizekey_body(long long b, const SizeKeyArgs &a) {
    a.keys[b] = ((u64)a.o_set[b] << 32) | (u32)(0x7FFFFFFF - a.size[b]);
    a.vals[b] = (u32)b;
}
MAP_KERNEL(sizekey, SizeKeyArgs, 20)

struct InvArgs { const u32 *order; u32 *inv; };
HD void inv_body(long long i, const InvArgs &a) { a.inv[a.order[i]] = (u32)i; }
MAP_KERNEL(inv, InvArgs, 8)

struct FinalArgs {
    BatchView v; const u32 *order; const u32 *inv; const u32 *o_depth; const u32 *o_set; const int *o_pos;
    const int *size; const int *total; const int *interval; const int *next;
    const u32 *set_blk0; const u32 *set_pos0;
    int *f_depth; int *f_size; int *f_total; int *f_interval; int *f_next; int *f_pos;
    const u32 *order0; const u32 *blk_leaf; u32 *f_leaf; // block b of the list order = block order0[b] of the emit order
};
HD void final_body(long long i, const FinalArgs &a) {
    u32 b = a.order[i];
    a.f_leaf[i] = a.blk_leaf[a.order0[b]];
    u32 s = a.o_set[b];
    u32 m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
    a.f_depth[i] = (int)a.o_depth[b];
    a.f_size[i] = a.size[b];
    a.f_total[i] = a.total[b];
    a.f_interval[i] = a.interval[b];
    int nx = a.next[b];
    a.f_next[i] = nx < 0 ? -1 : (int)(a.inv[nx] - LDG(a.set_blk0 + s));
    u32 src = pos_offset(a.set_blk0, a.set_pos0, s, m, b), dst = pos_offset(a.set_blk0, a.set_pos0, s, m, (u32)i);
    for (u32 k = 0; k < m; k++) a.f_pos[dst + k] = a.o_pos[src + k];
}
MAP_KERNEL(final, FinalArgs, 48)

// getRotations (csamsa.c:311): the positions of the head of the sorted list; plus whether the
// head chain bites its own tail (the reference then overruns blockLabel, nodeslinkedlists.c:161)
struct RotArgs {
    BatchView v; const u32 *set_blk0; const u32 *set_pos0; const int *f_pos; const int *f_next;
    int *rotations; u32 *set_cyclic;
};
HD void rot_body(long long s, const RotArgs &a) {
    u32 q0 = a.v.set_seq0[s], q1 = a.v.set_seq0[s + 1];
    u32 b0 = a.set_blk0[s], b1 = a.set_blk0[s + 1];
    a.set_cyclic[s] = 0;
    if (b0 == b1) { for (u32 k = q0; k < q1; k++) a.rotations[k] = 0; return; }
    u32 po = a.set_pos0[s];
    for (u32 k = q0; k < q1; k++) a.rotations[k] = a.f_pos[po + (k - q0)];
    u32 B = b1 - b0, steps = 0;
    int cur = 0;
    while (cur != -1 && steps <= B) { cur = a.f_next[b0 + cur]; steps++; }
    if (cur != -1) a.set_cyclic[s] = 1;
}
MAP_KERNEL(rot, RotArgs, 8)

// ---- the letters of a block as the reference prints them (nodeslinkedlists.c:128 blockLabel) -----------------
// blockLabel spells a block from the labels of the tree edges on its path, i.e. from the text that CREATED each
// edge (labelfrom/startpos, gencycsuffixtrees.c:160-218).  The edge that holds letter number j of block X was
// created by the first rotation, in insertion order, that begins with X[0..j]: a rotation of sequence 0 (every
// block occurs there), at the smallest such position p: letter = texts[0][p+j].  A, C, G, T are the same at every
// occurrence; a letter outside ACGT is spelled as that first occurrence has it.  p = the smallest position among
// the rotations of sequence 0 that share the block's first j+1 letters (seq0_child_minpos).
struct BlockLettersArgs {
    BatchView v; const unsigned char *raw; Seq0Q q;
    const int *f_depth; const int *f_pos; const u32 *f_leaf; const u32 *f_set; const u32 *set_blk0; const u32 *set_pos0;
    const unsigned long long *offsets; u32 B; char *out;
};
HD void blockletters_body(long long t, const BlockLettersArgs &a) {
    u32 lo = 0, hi = a.B; // block of letter t: last b with offsets[b] <= t
    while (lo < hi) { u32 mid = (lo + hi) >> 1; if (a.offsets[mid] <= (unsigned long long)t) lo = mid + 1; else hi = mid; }
    const u32 b = lo - 1, j = (u32)((unsigned long long)t - a.offsets[b]);
    const u32 s = a.f_set[b];
    const u32 k0 = LDG(a.v.set_seq0 + s), m = LDG(a.v.set_seq0 + s + 1) - k0;
    const u32 off = LDG(a.v.seq_off + k0), n0 = LDG(a.v.seq_off + k0 + 1) - off;
    const u32 p0 = (u32)a.f_pos[pos_offset(a.set_blk0, a.set_pos0, s, m, b)];
    unsigned char c = a.raw[off + (p0 + j) % n0];
    // (the rotations that begin with the block's first j+1 letters: the child, below depth j, that holds the block's leaf)
    if (code_of_letter(c) > 3) c = a.raw[off + (seq0_child_minpos(a.q, a.f_leaf[b], j) + j) % n0];
    a.out[t] = (char)c;
}
MAP_KERNEL(blockletters, BlockLettersArgs, 2)

// ---- optional: the counts the reference prints (csamsa.c:332 "nodes found", :338 "nodes left") ------------
// collectNodes keeps the DEEPEST nodes that hold every sequence.  With W(l) = [l, R[l]] the shortest
// window from l that holds every sequence and D(l) = min lcp over (l, R[l]] the depth of the
// smallest LCP interval around it: the windows inside one node are consecutive l, and a node is a
// deepest all-sequence node iff all of its windows have D == its depth.  So the collected nodes are
// exactly the maximal runs of equal D(l) whose two neighbouring windows are shallower (or missing).
struct WinDepthArgs { BatchView v; const u32 *sa; const u32 *lcp; const u32 *R; u32 *dv; };
HD void windepth_body(long long i, const WinDepthArgs &a) {
    u32 l = (u32)i;
    u32 s = set_of_pos(a.v, l);
    u32 s1 = LDG(a.v.set_base0 + s + 1);
    u32 r = a.R[l];
    if (r >= s1) { a.dv[l] = 0; return; } // no window: sorts below every depth
    u32 mn = 0xFFFFFFFEu;
    for (u32 j = l + 1; j <= r; j++) { u32 x = a.lcp[j]; if (x < mn) mn = x; }
    a.dv[l] = mn + 1;
}
MAP_KERNEL(windepth, WinDepthArgs, 12)

// (letter before the suffix, sequence) lists: prevcl[i] = 1 + previous SA index of the same sequence
// preceded by the same letter, 0 if none
struct ClKeyArgs { BatchView v; const u32 *sa; u64 *keys; u32 *vals; int mbits; };
HD unsigned letter_before(const BatchView &v, u32 g, u32 k) {
    u32 off = LDG(v.seq_off + k), n = LDG(v.seq_off + k + 1) - off;
    u32 p = g - off;
    return v.code[off + (p == 0 ? n - 1 : p - 1)];
}
HD void clkey_body(long long i, const ClKeyArgs &a) {
    u32 g = a.sa[i];
    u32 k = seq_of(a.v, g);
    u32 color = k - LDG(a.v.set_seq0 + LDG(a.v.seq_set + k));
    a.keys[i] = ((u64)letter_before(a.v, g, k) << a.mbits) | color;
    a.vals[i] = (u32)i;
}
MAP_KERNEL(clkey, ClKeyArgs, 24)

struct PrevClArgs { BatchView v; const u32 *sa; const u64 *keys; const u32 *vals; u32 *prevcl; int mbits; };
HD void prevcl_body(long long j, const PrevClArgs &a) {
    u32 i = a.vals[j];
    u32 p = 0;
    if (j > 0 && a.keys[j - 1] == a.keys[j]) {
        u32 i0 = a.vals[j - 1];
        if (seq_of(a.v, a.sa[i0]) == seq_of(a.v, a.sa[i])) p = i0 + 1;
    }
    a.prevcl[i] = p;
}
MAP_KERNEL(prevcl, PrevClArgs, 24)

struct PlateauArgs {
    BatchView v; const u32 *sa; const u32 *lcp; const u32 *R; const u32 *dv; const u32 *prevcl;
    u32 *set_collected; u32 *set_suffixfree;
};
HD void plateau_body(long long i, const PlateauArgs &a) {
    u32 l = (u32)i;
    u32 d1 = a.dv[l];
    if (d1 == 0) return;
    u32 k0 = seq_of(a.v, a.sa[l]);
    u32 s = LDG(a.v.seq_set + k0);
    u32 s0 = LDG(a.v.set_base0 + s), s1 = LDG(a.v.set_base0 + s + 1);
    if (l != s0) {
        u32 dl = a.dv[l - 1];
        if (dl >= d1) return; // not the first window of a run, or a deeper neighbour
    }
    u32 l2 = l;
    while (l2 + 1 < s1 && a.dv[l2 + 1] == d1) l2++;
    if (l2 + 1 < s1 && a.dv[l2 + 1] > d1) return;
    ATOMIC_ADD(a.set_collected + s, 1u);
    u32 d = d1 - 1;
    if (d == 0) { ATOMIC_ADD(a.set_suffixfree + s, 1u); return; } // csamsa.c:85
    u32 rb = a.R[l2];
    while (rb + 1 < s1 && a.lcp[rb + 1] >= d) rb++;
    // removeSuffixNodes (csamsa.c:80): some letter x precedes an occurrence in EVERY sequence
    u32 m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
    u32 cnt[5] = {0, 0, 0, 0, 0};
    for (u32 j = l; j <= rb; j++) {
        if (a.prevcl[j] > l) continue; // an earlier suffix of this sequence in the node has the same letter
        u32 g = a.sa[j];
        cnt[letter_before(a.v, g, seq_of(a.v, g))]++;
    }
    bool sfx = false;
    for (int c = 0; c < 5; c++) if (cnt[c] == m) sfx = true;
    if (!sfx) ATOMIC_ADD(a.set_suffixfree + s, 1u);
}
MAP_KERNEL(plateau, PlateauArgs, 16)

// ---- stage 1, fast path: one doubling round as ONE pass over the suffix array ---------------------------
// After the first sort the groups of equal h-prefix are small (about one suffix per sequence and
// repeat copy), so a round does not need a device-wide sort: the array is cut into tiles that begin
// at group borders (<= RF_CAP suffixes), each CTA stages its tile in shared memory, ranks every
// suffix inside its own group by (rank of the next h letters, position) -- a stable counting rank --
// and writes the tile back with the new group borders and ranks.  HBM traffic per suffix and round:
// sa 4 B in + 4 B out, head 4 + 4, the two gathers seqof/rank 4 + 4, the new rank 4 = 28 B (the
// device-wide radix path moves 7 x 32 B).  Ranks are double buffered: a round reads the ranks of
// the previous round only.  A group larger than a tile sends the round down the device-wide path.
#define RF_THREADS 256
#define RF_NOMINAL 1024
#define RF_CAP 2048
#define RF_ITEMS (RF_CAP / RF_THREADS)
#define RF_QUAD_GROUP 128u // quadrupling rounds (k_refine4) pay off while the groups are this small
#define RF_WARP_GROUP 512u // groups up to this size are ranked by one warp (min-reductions); longer ones by counting

struct TileArgs { const u32 *head; u32 *tb; u32 *oversize; u32 N; u32 ntiles; };
HD u32 tile_start(const TileArgs &a, u32 t) {
    if (t >= a.ntiles) return a.N;
    u64 p = (u64)t * RF_NOMINAL;
    u32 steps = 0;
    while (p < a.N && (a.head[p] & 0x7FFFFFFFu) != (u32)p) { // bit 31: settled (k_refine)
        p++;
        if (++steps > RF_CAP) break;
    }
    return p < a.N ? (u32)p : a.N;
}
HD void tile_body(long long t, const TileArgs &a) {
    u32 b0 = tile_start(a, (u32)t), b1 = tile_start(a, (u32)t + 1);
    a.tb[t] = b0;
    if ((u32)t + 1 == a.ntiles) a.tb[t + 1] = a.N;
    if (b1 - b0 > RF_CAP) ATOMIC_MAX(a.oversize, 1u);
}
MAP_KERNEL(tile, TileArgs, 8)

// largest group of equal h-prefixes: the last suffix of a group is as far from its head as the group is long
// (pairs != nullptr: also pairs[0] = the number of pairs of suffixes that share a group -- what the word sort would
// compare --, pairs[1] = the number of suffixes that share a group and pairs[2] = those of them whose group holds no more
// than a warp has lanes, which the carried word sort can walk; pairs[3] = those in groups of no more than 256, which it walks
// with a CTA per group)
struct MaxGroupArgs { const u32 *head; u32 *maxgroup; u32 N; unsigned long long *pairs; };
#ifdef CSA_EMU
HD void maxgroup_body(long long i, const MaxGroupArgs &a) {
    if ((u32)i + 1 == a.N || (a.head[i + 1] & 0x7FFFFFFFu) == (u32)i + 1) {
        u32 sz = (u32)i - (a.head[i] & 0x7FFFFFFFu) + 1;
        if (sz > *a.maxgroup) *a.maxgroup = sz;
        if (a.pairs) { a.pairs[0] += (unsigned long long)sz * (sz - 1) / 2; if (sz > 1) a.pairs[1] += sz; if (sz > 1 && sz <= 32) a.pairs[2] += sz; if (sz > 1 && sz <= 256) a.pairs[3] += sz; }
    }
}
MAP_KERNEL(maxgroup, MaxGroupArgs, 4)
#else
__global__ void __launch_bounds__(256) k_maxgroup(long long n, MaxGroupArgs a) {
    __shared__ u32 s_max;
    __shared__ unsigned long long s_pairs, s_shared, s_walk, s_walk2;
    if (threadIdx.x == 0) { s_max = 0; s_pairs = 0; s_shared = 0; s_walk = 0; s_walk2 = 0; }
    __syncthreads();
    // a persistent grid: every thread folds many places (neighbouring lanes read neighbouring words)
    u32 sz = 0, sh = 0, wk = 0, wk2 = 0;
    unsigned long long pr = 0;
    // (four places a round, their heads loaded before any is looked at: the loads of a round are in flight together)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        u32 hn[4], hc[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const long long i = i0 + q * stride;
            hn[q] = (i < n && (u32)i + 1 != a.N) ? a.head[i + 1] & 0x7FFFFFFFu : 0xFFFFFFFFu;
            hc[q] = i < n ? a.head[i] & 0x7FFFFFFFu : 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const long long i = i0 + q * stride;
            if (i < n && ((u32)i + 1 == a.N || hn[q] == (u32)i + 1)) {
                const u32 z = (u32)i - hc[q] + 1;
                sz = z > sz ? z : sz;
                if (z > 1) { pr += (unsigned long long)z * (z - 1) / 2; sh += z; if (z <= 32u) wk += z; if (z <= 256u) wk2 += z; }
            }
        }
    }
    if (a.pairs) { // one atomic per CTA: same-address atomics serialise in L2
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { pr += __shfl_xor_sync(0xffffffffu, pr, d); sh += __shfl_xor_sync(0xffffffffu, sh, d); wk += __shfl_xor_sync(0xffffffffu, wk, d); wk2 += __shfl_xor_sync(0xffffffffu, wk2, d); }
        if ((threadIdx.x & 31) == 0 && pr) { atomicAdd(&s_pairs, pr); atomicAdd(&s_shared, (unsigned long long)sh); atomicAdd(&s_walk, (unsigned long long)wk); atomicAdd(&s_walk2, (unsigned long long)wk2); }
    }
    sz = __reduce_max_sync(0xffffffffu, sz);
    if ((threadIdx.x & 31) == 0 && sz) atomicMax(&s_max, sz);
    __syncthreads();
    if (threadIdx.x == 0 && s_max) atomicMax(a.maxgroup, s_max);
    if (threadIdx.x == 0 && a.pairs && s_pairs) { atomicAdd(a.pairs, s_pairs); atomicAdd(a.pairs + 1, s_shared); atomicAdd(a.pairs + 2, s_walk); atomicAdd(a.pairs + 3, s_walk2); }
}
static inline void launch_maxgroup(Exec &ex, long long n, MaxGroupArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_maxgroup", 4.0 * n);
    const long long want = (n + 255) / 256, cap = 148 * 16;
    k_maxgroup<<<(unsigned)(want < cap ? want : cap), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

struct RefineArgs {
    BatchView v; u32 *sa; u32 *head; const u32 *rank; u32 *rank2; const u32 *tb; u32 h; u32 *ngroups; u32 ntiles;
    u32 *maxgroup; // largest group after the round (atomic max)
    u32 *staged;   // suffixes that were not settled yet, i.e. really read, ranked and written (for the profile)
    u32 *active;   // suffixes that still share a group after the round (warp paths only; else = staged)
};

#ifdef CSA_EMU
static inline void emu_refine(const RefineArgs &a, int nkeys) {
    struct Item { u32 k[3]; u32 g; };
    std::vector<Item> seg;
    for (u32 t = 0; t < a.ntiles; t++) {
        u32 base = a.tb[t], end = a.tb[t + 1];
        u32 i = base;
        while (i < end) {
            u32 j = i + 1;
            while (j < end && (a.head[j] & 0x7FFFFFFFu) != j) j++;
            seg.clear();
            for (u32 x = i; x < j; x++) {
                Item it{{0, 0, 0}, a.sa[x]};
                u32 g = a.sa[x];
                for (int q = 0; q < nkeys; q++) { g = cyc_add(a.v, g, a.h); it.k[q] = a.rank[g]; }
                seg.push_back(it);
            }
            auto less = [](const Item &p, const Item &q) {
                for (int w = 0; w < 3; w++) if (p.k[w] != q.k[w]) return p.k[w] < q.k[w];
                return false;
            };
            std::stable_sort(seg.begin(), seg.end(), less);
            u32 hd = i;
            for (u32 x = i; x < j; x++) {
                if (x > i && less(seg[x - i - 1], seg[x - i])) hd = x;
                if (hd == x) (*a.ngroups)++;
                if (x - hd + 1 == 2) *a.active += 2; else if (x - hd + 1 > 2) *a.active += 1; // suffixes still sharing a group
                if (x - hd + 1 > *a.maxgroup) *a.maxgroup = x - hd + 1;
                a.sa[x] = seg[x - i].g;
                a.head[x] = hd;
                a.rank2[seg[x - i].g] = hd;
            }
            i = j;
        }
    }
}
static inline void launch_refine(Exec &, const RefineArgs &a) { emu_refine(a, 1); }
static inline void launch_refine4(Exec &, const RefineArgs &a) { emu_refine(a, 3); }
#else
#define HEAD_SETTLED 0x80000000u
__global__ void __launch_bounds__(RF_THREADS, 4) k_refine(RefineArgs a) {
    // composite sort key of a suffix: (first place of its group in the tile : 11 bits, rank of the
    // next h letters : 31, place in the tile : 11).  It is unique and orders the whole tile, so the
    // number of smaller composites from the first group a thread touches onwards IS the sorted
    // place -- the counting loop needs no test for group membership.
    __shared__ u64 s_ck[RF_CAP];
    __shared__ u32 s_sa[RF_CAP];
    __shared__ u32 s_k2[RF_CAP];   // rank h letters on | 1<<31 on the first suffix of a group; later the new heads
    __shared__ u32 s_osa[RF_CAP];
    __shared__ unsigned char s_fl[RF_CAP + 4]; // bit 0: first of its group, bit 1: settled, bit 2 (at a group's
                                               // first place): some member's second rank differs
    __shared__ u32 s_scan[33];
    __shared__ u32 s_count, s_big, s_mid, s_maxg, s_staged, s_act;
    u32 *s_okey = (u32 *)s_ck;     // sorted keys; s_ck is dead (and fenced) by then
    const u32 base = a.tb[blockIdx.x];
    const u32 n = a.tb[blockIdx.x + 1] - base;
    if (n == 0 || n > RF_CAP) return;
    const u32 tid = threadIdx.x;
    if (tid == 0) { s_count = 0; s_big = 0; s_mid = 0; s_maxg = 1; s_staged = 0; s_act = 0; }
    // 1. borders.  A settled suffix is a singleton whose rank is already in BOTH rank buffers.
    //    (every phase that loads from HBM issues all RF_ITEMS loads of a thread before using any:
    //    the gathers below are chains of four dependent L2/HBM accesses and need the overlap)
    int all_settled = 1;
    {
        u32 hdv[RF_ITEMS];
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            hdv[x] = (j < n) ? a.head[base + j] : 0u;
        }
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            if (j < n) {
                u32 fl = ((hdv[x] & ~HEAD_SETTLED) == base + j ? 1u : 0u) | ((hdv[x] >> 31) << 1);
                s_fl[j] = (unsigned char)fl;
                all_settled &= (int)(fl >> 1);
            }
        }
    }
    if (__syncthreads_and(all_settled)) { // every suffix of the tile has its final place
        if (tid == 0) { atomicAdd(a.ngroups, n); atomicMax(a.maxgroup, 1u); }
        return;
    }
    // 2. stage the suffixes; only those that still share a group pay for the gathers
    {
        u32 gv[RF_ITEMS], kv[RF_ITEMS], ov[RF_ITEMS], nv[RF_ITEMS], k2v[RF_ITEMS];
        unsigned need = 0, share = 0; // bit x: my x-th suffix must be staged / still shares its group
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            if (j < n) {
                u32 fl = s_fl[j];
                bool single = (fl & 1u) && (j + 1 == n || (s_fl[j + 1] & 1u));
                if (!(single && (fl & 2u))) need |= 1u << x;
                if (!single) share |= 1u << x;
            }
        }
        if (need) atomicAdd(&s_staged, (u32)__popc(need));
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) gv[x] = (need >> x & 1u) ? a.sa[base + tid + x * RF_THREADS] : 0u;
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) kv[x] = (share >> x & 1u) ? seq_of(a.v, gv[x]) : 0u;
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            ov[x] = (share >> x & 1u) ? LDG(a.v.seq_off + kv[x]) : 0u;
            nv[x] = (share >> x & 1u) ? LDG(a.v.seq_off + kv[x] + 1) : 1u;
        }
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            k2v[x] = 0;
            if (share >> x & 1u) {
                u32 len = nv[x] - ov[x], hh = a.h;
                if (hh >= len) hh %= len;
                u32 q = gv[x] - ov[x] + hh;
                if (q >= len) q -= len;
                k2v[x] = LDG(a.rank + ov[x] + q);
            }
        }
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            if (j < n) {
                if (need >> x & 1u) s_sa[j] = gv[x];
                s_k2[j] = k2v[x] | ((u32)(s_fl[j] & 1u) << 31);
            }
        }
    }
    __syncthreads();
    // 3. first place of the group of each of my RF_ITEMS consecutive suffixes (block max-scan), and the
    //    list of the groups that still hold more than one suffix (block sum-scan)
    const u32 j0 = tid * RF_ITEMS;
    u64 myck[RF_ITEMS];
    u32 seg[RF_ITEMS];
    u32 run = 0, nlist = 0, listmask = 0;
    bool lonely = true; // all of mine stand alone: nothing to count
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) {
        u32 j = j0 + e;
        u32 w = (j < n) ? s_k2[j] : 0x80000000u;
        if (w >> 31) run = j;
        seg[e] = run;
        myck[e] = (u64)(w & 0x7FFFFFFFu);
        if (j < n && !((w >> 31) && (j + 1 == n || (s_k2[j + 1] >> 31)))) {
            lonely = false;
            if (w >> 31) { nlist++; listmask |= 1u << e; }
        }
    }
    u32 total;
    u32 before = block_scan_excl(run, total, ScanMax(), s_scan);
    bool big = false, mid = false;
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) {
        u32 j = j0 + e;
        seg[e] = seg[e] > before ? seg[e] : before;
        if (j < n && j - seg[e] >= RF_WARP_GROUP) big = true;
        if (j < n && j - seg[e] >= 32u) mid = true;
    }
    if (big) s_big = 1; // same value from every writer; read after the barriers of the scan below
    if (mid) s_mid = 1;
    u32 ngroups_listed;
    u32 lbefore = block_scan_excl(nlist, ngroups_listed, ScanSum(), s_scan);
    if (!s_big) {
        // 3w. no group of the tile is longer than a warp: one warp per group, one suffix per lane.
        //     Ranks come from repeated min-reductions over the group's second ranks: every distinct
        //     value is one new group, its members keep their order (stable), and the new head falls
        //     out of the same loop -- no composite keys, no counting loop, no second scan.
        u32 *s_gstart = s_osa;
#pragma unroll
        for (int e = 0; e < RF_ITEMS; e++)
            if (listmask >> e & 1u) s_gstart[lbefore++] = j0 + e;
        __syncthreads();
        const unsigned lane = tid & 31u, warp = tid >> 5, ltmask = (1u << lane) - 1u;
        u32 made = 0, widest = 0, sharing = 0;
        for (u32 gi = warp; gi < ngroups_listed; gi += RF_THREADS / 32) {
            const u32 S = s_gstart[gi];
            // the group ends where the next one begins (a border flag; the end of the tile counts as one)
            u32 size = 0;
            if (!s_mid) { // no group of the tile is longer than a warp: the next border is within reach
                const u32 j = S + lane;
                const u32 w = (j < n) ? s_k2[j] : 0x80000000u;
                const unsigned nextstart = __ballot_sync(0xffffffffu, lane > 0 && (w >> 31));
                size = nextstart ? (u32)(__ffs((int)nextstart) - 1) : 32u;
            } else for (u32 c0 = 0;; c0 += 32) {
                const u32 j = S + c0 + lane;
                const u32 w = (j < n) ? s_k2[j] : 0x80000000u;
                const unsigned nextstart = __ballot_sync(0xffffffffu, (c0 + lane > 0) && (w >> 31));
                if (nextstart) { size = c0 + (u32)(__ffs((int)nextstart) - 1); break; }
            }
            if (size <= 32) { // one suffix per lane, everything in registers
                const u32 j = S + lane;
                const bool member = lane < size;
                const u32 k2 = member ? (s_k2[j] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
                const u32 g = member ? s_sa[j] : 0u;
                unsigned rem = __ballot_sync(0xffffffffu, member);
                u32 placed = 0, mynew = 0, myhead = 0;
                while (rem) {
                    const u32 kk = (rem >> lane & 1u) ? k2 : 0xFFFFFFFFu;
                    const u32 mn = __reduce_min_sync(0xffffffffu, kk);
                    const unsigned eq = __ballot_sync(0xffffffffu, kk == mn) & rem;
                    if (eq >> lane & 1u) { mynew = placed + __popc(eq & ltmask); myhead = placed; }
                    const u32 c = (u32)__popc(eq);
                    placed += c;
                    widest = c > widest ? c : widest;
                    if (c > 1) sharing += c;
                    rem &= ~eq;
                    made++;
                }
                if (member) {
                    const u32 p = base + S + mynew, h2 = base + S + myhead;
                    a.sa[p] = g;
                    a.head[p] = h2;
                    a.rank2[g] = h2;
                }
            } else { // several suffixes per lane (striped), second ranks stay in shared memory;
                     // a placed suffix is marked by an all-ones word (its border bit is never read again)
                u32 placed = 0;
                while (placed < size) {
                    u32 mn = 0xFFFFFFFFu;
                    for (u32 t = lane; t < size; t += 32) {
                        u32 w = s_k2[S + t];
                        if (w != 0xFFFFFFFFu) { w &= 0x7FFFFFFFu; mn = w < mn ? w : mn; }
                    }
                    mn = __reduce_min_sync(0xffffffffu, mn);
                    u32 cnt = 0;
                    for (u32 t0 = 0; t0 < size; t0 += 32) {
                        const u32 t = t0 + lane;
                        u32 w = (t < size) ? s_k2[S + t] : 0xFFFFFFFFu;
                        const bool eq = w != 0xFFFFFFFFu && (w & 0x7FFFFFFFu) == mn;
                        const unsigned b = __ballot_sync(0xffffffffu, eq);
                        if (eq) {
                            const u32 p = base + S + placed + cnt + __popc(b & ltmask), h2 = base + S + placed;
                            const u32 g = s_sa[S + t];
                            a.sa[p] = g;
                            a.head[p] = h2;
                            a.rank2[g] = h2;
                            s_k2[S + t] = 0xFFFFFFFFu;
                        }
                        cnt += __popc(b);
                    }
                    placed += cnt;
                    widest = cnt > widest ? cnt : widest;
                    if (cnt > 1) sharing += cnt;
                    made++;
                }
            }
        }
        if (lane == 0 && widest > 1) atomicMax(&s_maxg, widest);
        if (lane == 0 && sharing) atomicAdd(&s_act, sharing);
        u32 nsingle = 0;
        for (u32 j = tid; j < n; j += RF_THREADS) {
            u32 fl = s_fl[j];
            if ((fl & 1u) && (j + 1 == n || (s_fl[j + 1] & 1u))) {
                nsingle++;
                if (!(fl & 2u)) {
                    a.rank2[s_sa[j]] = base + j;
                    a.head[base + j] = (base + j) | HEAD_SETTLED;
                }
            }
        }
        if (lane == 0) nsingle += made;
        if (nsingle) atomicAdd(&s_count, nsingle);
        __syncthreads();
        if (tid == 0) { atomicAdd(a.ngroups, s_count); atomicMax(a.maxgroup, s_maxg); atomicAdd(a.staged, s_staged); if (s_act) atomicAdd(a.active, s_act); }
        return;
    }
    // 3c. a group longer than a warp: composite keys and a counting rank over shared memory
    u32 seg_last = 0;
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) {
        u32 j = j0 + e;
        myck[e] = ((u64)seg[e] << 42) | (myck[e] << 11) | j;
        if (j < n) {
            s_ck[j] = myck[e];
            seg_last = seg[e];
            // a group whose members all carry the same second rank keeps its order: nothing to count
            if ((u32)(myck[e] >> 11 & 0x7FFFFFFFu) != (s_k2[seg[e]] & 0x7FFFFFFFu)) s_fl[seg[e]] |= 4; // same bit from every writer
        }
    }
    __syncthreads();
    // 4. sorted place = first place of my first group + number of smaller composites from there on
    u32 cnt[RF_ITEMS];
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) cnt[e] = j0 + e - seg[0];
    if (!lonely) {
        lonely = true;
#pragma unroll
        for (int e = 0; e < RF_ITEMS; e++)
            if (j0 + e < n && (s_fl[seg[e]] & 4)) lonely = false;
    }
    if (!lonely) {
#pragma unroll
        for (int e = 0; e < RF_ITEMS; e++) cnt[e] = 0;
        for (u32 k = seg[0]; k < n; k++) {
            u64 ck = s_ck[k];
            if ((u32)(ck >> 42) > seg_last) break;
#pragma unroll
            for (int e = 0; e < RF_ITEMS; e++) cnt[e] += (ck < myck[e]) ? 1u : 0u;
        }
    }
    __syncthreads(); // s_ck is dead from here: s_okey may overwrite it
    // 5. move; the suffix that lands on its group's first place carries the border
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) {
        u32 j = j0 + e;
        if (j < n) {
            u32 p = seg[0] + cnt[e];
            s_osa[p] = s_sa[j];
            s_okey[p] = ((u32)(myck[e] >> 11) & 0x7FFFFFFFu) | (p == seg[e] ? 0x80000000u : 0u);
        }
    }
    __syncthreads();
    // 6. new borders: an old border, or a change of the second rank; new heads by max-scan
    u32 hd[RF_ITEMS];
    u32 run2 = 0, nflag = 0;
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) {
        u32 j = j0 + e;
        if (j < n) {
            u32 w = s_okey[j];
            bool f = (w >> 31) || j == 0 || (w & 0x7FFFFFFFu) != (s_okey[j - 1] & 0x7FFFFFFFu);
            if (f) { run2 = j; nflag++; }
        }
        hd[e] = run2;
    }
    u32 total2;
    u32 before2 = block_scan_excl(run2, total2, ScanMax(), s_scan);
    if (nflag) atomicAdd(&s_count, nflag);
    u32 widest = 1;
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) {
        u32 j = j0 + e;
        if (j < n) {
            u32 hl = hd[e] > before2 ? hd[e] : before2;
            s_k2[j] = base + hl;
            widest = (j - hl + 1 > widest) ? j - hl + 1 : widest; // the last suffix of a group tells its size
        }
    }
    if (widest > 1) atomicMax(&s_maxg, widest);
    if (tid == 0) s_act = n; // upper bound: this path does not count the suffixes still sharing
    __syncthreads();
    // 7. write back, coalesced.  The new rank goes to the OTHER rank buffer (a round reads only the
    //    ranks of the round before); a suffix that stood alone at the start keeps its place and is
    //    settled from now on: its rank is in both buffers.
    for (u32 j = tid; j < n; j += RF_THREADS) {
        u32 fl = s_fl[j];
        bool single = (fl & 1u) && (j + 1 == n || (s_fl[j + 1] & 1u));
        if (single) {
            if (!(fl & 2u)) {
                a.rank2[s_osa[j]] = base + j;
                a.head[base + j] = (base + j) | HEAD_SETTLED;
            }
        } else {
            u32 g = s_osa[j], h2 = s_k2[j];
            a.sa[base + j] = g;
            a.head[base + j] = h2;
            a.rank2[g] = h2;
        }
    }
    if (tid == 0) { atomicAdd(a.ngroups, s_count); atomicMax(a.maxgroup, s_maxg); atomicAdd(a.staged, s_staged); if (s_act) atomicAdd(a.active, s_act); }
}

// ---- the same round, four times the letters: groups ordered by the ranks h, 2h and 3h letters on ----------
// (prefix "quadrupling": half as many rounds, each with three rank gathers instead of one).  One warp
// per group, lexicographic minimum of the rank triple by three chained min-reductions; any group
// size works, the host picks this kernel when the groups are small enough to be quick (<= 512).
__global__ void __launch_bounds__(RF_THREADS, 4) k_refine4(RefineArgs a) {
    __shared__ u32 s_sa[RF_CAP];
    __shared__ u32 s_ka[RF_CAP]; // rank h on | 1<<31 on the first suffix of a group; all ones once placed
    __shared__ u32 s_kb[RF_CAP], s_kc[RF_CAP];
    __shared__ u32 s_gstart[RF_CAP / 2 + 1];
    __shared__ unsigned char s_fl[RF_CAP + 4]; // bit 0: first of its group, bit 1: settled
    __shared__ u32 s_scan[33];
    __shared__ u32 s_count, s_maxg, s_staged, s_act;
    const u32 base = a.tb[blockIdx.x];
    const u32 n = a.tb[blockIdx.x + 1] - base;
    if (n == 0 || n > RF_CAP) return;
    const u32 tid = threadIdx.x;
    if (tid == 0) { s_count = 0; s_maxg = 1; s_staged = 0; s_act = 0; }
    int all_settled = 1;
    {
        u32 hdv[RF_ITEMS];
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            hdv[x] = (j < n) ? a.head[base + j] : 0u;
        }
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            if (j < n) {
                u32 fl = ((hdv[x] & ~HEAD_SETTLED) == base + j ? 1u : 0u) | ((hdv[x] >> 31) << 1);
                s_fl[j] = (unsigned char)fl;
                all_settled &= (int)(fl >> 1);
            }
        }
    }
    if (__syncthreads_and(all_settled)) {
        if (tid == 0) { atomicAdd(a.ngroups, n); atomicMax(a.maxgroup, 1u); }
        return;
    }
    {
        u32 gv[RF_ITEMS], kv[RF_ITEMS], ov[RF_ITEMS], nv[RF_ITEMS];
        unsigned need = 0, share = 0;
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            if (j < n) {
                u32 fl = s_fl[j];
                bool single = (fl & 1u) && (j + 1 == n || (s_fl[j + 1] & 1u));
                if (!(single && (fl & 2u))) need |= 1u << x;
                if (!single) share |= 1u << x;
            }
        }
        if (need) atomicAdd(&s_staged, (u32)__popc(need));
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) gv[x] = (need >> x & 1u) ? a.sa[base + tid + x * RF_THREADS] : 0u;
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) kv[x] = (share >> x & 1u) ? seq_of(a.v, gv[x]) : 0u;
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            ov[x] = (share >> x & 1u) ? LDG(a.v.seq_off + kv[x]) : 0u;
            nv[x] = (share >> x & 1u) ? LDG(a.v.seq_off + kv[x] + 1) : 1u;
        }
#pragma unroll
        for (int x = 0; x < RF_ITEMS; x++) {
            u32 j = tid + x * RF_THREADS;
            u32 ka = 0, kb = 0, kc = 0;
            if (share >> x & 1u) {
                u32 len = nv[x] - ov[x], hh = a.h;
                if (hh >= len) hh %= len;
                u32 q1 = gv[x] - ov[x] + hh; if (q1 >= len) q1 -= len;
                u32 q2 = q1 + hh; if (q2 >= len) q2 -= len;
                u32 q3 = q2 + hh; if (q3 >= len) q3 -= len;
                ka = LDG(a.rank + ov[x] + q1);
                kb = LDG(a.rank + ov[x] + q2);
                kc = LDG(a.rank + ov[x] + q3);
            }
            if (j < n) {
                if (need >> x & 1u) s_sa[j] = gv[x];
                s_ka[j] = ka | ((u32)(s_fl[j] & 1u) << 31);
                s_kb[j] = kb;
                s_kc[j] = kc;
            }
        }
    }
    __syncthreads();
    // the groups that still hold more than one suffix
    const u32 j0 = tid * RF_ITEMS;
    u32 nlist = 0, listmask = 0;
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++) {
        u32 j = j0 + e;
        if (j < n && (s_ka[j] >> 31) && !(j + 1 == n || (s_ka[j + 1] >> 31))) { nlist++; listmask |= 1u << e; }
    }
    u32 ngroups_listed;
    u32 lbefore = block_scan_excl(nlist, ngroups_listed, ScanSum(), s_scan);
#pragma unroll
    for (int e = 0; e < RF_ITEMS; e++)
        if (listmask >> e & 1u) s_gstart[lbefore++] = j0 + e;
    __syncthreads();
    const unsigned lane = tid & 31u, warp = tid >> 5, ltmask = (1u << lane) - 1u;
    u32 made = 0, widest = 0, sharing = 0;
    for (u32 gi = warp; gi < ngroups_listed; gi += RF_THREADS / 32) {
        const u32 S = s_gstart[gi];
        u32 size = 0;
        for (u32 c0 = 0;; c0 += 32) {
            const u32 j = S + c0 + lane;
            const u32 w = (j < n) ? s_ka[j] : 0x80000000u;
            const unsigned nextstart = __ballot_sync(0xffffffffu, (c0 + lane > 0) && (w >> 31));
            if (nextstart) { size = c0 + (u32)(__ffs((int)nextstart) - 1); break; }
        }
        if (size <= 32) {
            const u32 j = S + lane;
            const bool member = lane < size;
            const u32 ka = member ? (s_ka[j] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
            const u32 kb = member ? s_kb[j] : 0xFFFFFFFFu, kc = member ? s_kc[j] : 0xFFFFFFFFu;
            const u32 g = member ? s_sa[j] : 0u;
            unsigned rem = __ballot_sync(0xffffffffu, member);
            u32 placed = 0, mynew = 0, myhead = 0;
            while (rem) {
                const u32 xa = (rem >> lane & 1u) ? ka : 0xFFFFFFFFu;
                const u32 ma = __reduce_min_sync(0xffffffffu, xa);
                const unsigned m1 = __ballot_sync(0xffffffffu, xa == ma) & rem;
                unsigned eq = m1;
                if (m1 & (m1 - 1)) { // more than one suffix shares the smallest first rank: look further
                    const u32 xb = (m1 >> lane & 1u) ? kb : 0xFFFFFFFFu;
                    const u32 mb = __reduce_min_sync(0xffffffffu, xb);
                    const unsigned m2 = __ballot_sync(0xffffffffu, xb == mb) & m1;
                    eq = m2;
                    if (m2 & (m2 - 1)) {
                        const u32 xc = (m2 >> lane & 1u) ? kc : 0xFFFFFFFFu;
                        const u32 mc = __reduce_min_sync(0xffffffffu, xc);
                        eq = __ballot_sync(0xffffffffu, xc == mc) & m2;
                    }
                }
                if (eq >> lane & 1u) { mynew = placed + __popc(eq & ltmask); myhead = placed; }
                const u32 c = (u32)__popc(eq);
                placed += c;
                widest = c > widest ? c : widest;
                if (c > 1) sharing += c;
                rem &= ~eq;
                made++;
            }
            if (member) {
                const u32 p = base + S + mynew, h2 = base + S + myhead;
                a.sa[p] = g;
                a.head[p] = h2;
                a.rank2[g] = h2;
            }
        } else {
            u32 placed = 0;
            while (placed < size) {
                u32 ma = 0xFFFFFFFFu, mb = 0xFFFFFFFFu, mc = 0xFFFFFFFFu;
                for (u32 t = lane; t < size; t += 32) {
                    u32 w = s_ka[S + t];
                    if (w != 0xFFFFFFFFu) { w &= 0x7FFFFFFFu; ma = w < ma ? w : ma; }
                }
                ma = __reduce_min_sync(0xffffffffu, ma);
                for (u32 t = lane; t < size; t += 32) {
                    u32 w = s_ka[S + t];
                    if (w != 0xFFFFFFFFu && (w & 0x7FFFFFFFu) == ma) { u32 y = s_kb[S + t]; mb = y < mb ? y : mb; }
                }
                mb = __reduce_min_sync(0xffffffffu, mb);
                for (u32 t = lane; t < size; t += 32) {
                    u32 w = s_ka[S + t];
                    if (w != 0xFFFFFFFFu && (w & 0x7FFFFFFFu) == ma && s_kb[S + t] == mb) { u32 y = s_kc[S + t]; mc = y < mc ? y : mc; }
                }
                mc = __reduce_min_sync(0xffffffffu, mc);
                u32 cnt = 0;
                for (u32 t0 = 0; t0 < size; t0 += 32) {
                    const u32 t = t0 + lane;
                    const u32 w = (t < size) ? s_ka[S + t] : 0xFFFFFFFFu;
                    const bool eq = w != 0xFFFFFFFFu && (w & 0x7FFFFFFFu) == ma && s_kb[S + t] == mb && s_kc[S + t] == mc;
                    const unsigned b = __ballot_sync(0xffffffffu, eq);
                    if (eq) {
                        const u32 p = base + S + placed + cnt + __popc(b & ltmask), h2 = base + S + placed;
                        const u32 g = s_sa[S + t];
                        a.sa[p] = g;
                        a.head[p] = h2;
                        a.rank2[g] = h2;
                        s_ka[S + t] = 0xFFFFFFFFu;
                    }
                    cnt += __popc(b);
                }
                placed += cnt;
                widest = cnt > widest ? cnt : widest;
                if (cnt > 1) sharing += cnt;
                made++;
            }
        }
    }
    if (lane == 0 && widest > 1) atomicMax(&s_maxg, widest);
    if (lane == 0 && sharing) atomicAdd(&s_act, sharing);
    u32 nsingle = 0;
    for (u32 j = tid; j < n; j += RF_THREADS) {
        u32 fl = s_fl[j];
        if ((fl & 1u) && (j + 1 == n || (s_fl[j + 1] & 1u))) {
            nsingle++;
            if (!(fl & 2u)) {
                a.rank2[s_sa[j]] = base + j;
                a.head[base + j] = (base + j) | HEAD_SETTLED;
            }
        }
    }
    if (lane == 0) nsingle += made;
    if (nsingle) atomicAdd(&s_count, nsingle);
    __syncthreads();
    if (tid == 0) { atomicAdd(a.ngroups, s_count); atomicMax(a.maxgroup, s_maxg); atomicAdd(a.staged, s_staged); if (s_act) atomicAdd(a.active, s_act); }
}
static inline void launch_refine4(Exec &ex, const RefineArgs &a) {
    if (a.ntiles == 0) return;
    PROF_BEGIN(ex, "k_refine4", 36.0 * a.v.N);
    k_refine4<<<a.ntiles, RF_THREADS, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
static inline void launch_refine(Exec &ex, const RefineArgs &a) {
    if (a.ntiles == 0) return;
    PROF_BEGIN(ex, "k_refine", 28.0 * a.v.N);
    k_refine<<<a.ntiles, RF_THREADS, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// ---- stage 1, group lists: a round touches only the suffixes that still share a group -----------------
// The tile rounds above read the head of EVERY suffix every round.  Once most suffixes stand alone
// that is wasted work, so the rounds can also run from a compact list of the groups that still hold
// two or more suffixes (start : 32 | size : 32).  One warp takes one group: its suffixes sit in
// consecutive places of the suffix array (one coalesced load), the ranks h (2h, 3h) letters on are
// gathered, the group is ranked by chained min-reductions as in k_refine4, and sa / head / the other
// rank buffer are written in place.  New groups of two or more go to the next round's list, new
// singletons to a list whose ranks are copied into the second rank buffer at the start of the next
// round (a round may only read ranks of the round before).  Lists are appended through shared
// memory: one atomic per CTA, not per group.  head[] stays in step, so tile rounds, device-wide
// rounds and list rounds can follow one another freely.
struct GListBuildArgs { const u32 *head; u32 N; u64 *list; u32 *count; };
struct RefineGArgs {
    BatchView v; u32 *sa; u32 *head; const u32 *rank; u32 *rank2; u32 h; int nkeys;
    const u64 *list; u32 nlist;
    u64 *next; u32 *n_next;       // groups of two or more made by this round
    u32 *singles; u32 *n_singles; // suffixes that came to stand alone in this round
    u32 *maxgroup; u32 *staged;
    u32 *ka, *kb, *kc, *gs;       // scratch [N], used by groups longer than a warp
};
struct CopySinglesArgs { const u32 *singles; const u32 *rank; u32 *rank2; u32 *head; };
HD void copysingles_body(long long i, const CopySinglesArgs &a) {
    u32 g = a.singles[i];
    u32 r = a.rank[g];
    a.rank2[g] = r;
    a.head[r] = r | 0x80000000u; // settled: alone, and its rank is in both buffers
}
MAP_KERNEL(copysingles, CopySinglesArgs, 16)

#ifdef CSA_EMU
static inline void launch_glist_build(Exec &, const GListBuildArgs &a) {
    for (u32 i = 0; i < a.N; i++)
        if (i + 1 == a.N || (a.head[i + 1] & 0x7FFFFFFFu) == i + 1) {
            u32 st = a.head[i] & 0x7FFFFFFFu;
            if (i - st + 1 >= 2) a.list[(*a.count)++] = ((u64)st << 32) | (i - st + 1);
        }
}
static inline void launch_refine_g(Exec &, const RefineGArgs &a) {
    struct Item { u32 k[3]; u32 g; };
    std::vector<Item> seg;
    for (u32 gi = 0; gi < a.nlist; gi++) {
        u32 start = (u32)(a.list[gi] >> 32), size = (u32)a.list[gi];
        seg.clear();
        for (u32 x = 0; x < size; x++) {
            Item it{{0, 0, 0}, a.sa[start + x]};
            u32 g = it.g;
            for (int q = 0; q < a.nkeys; q++) { g = cyc_add(a.v, g, a.h); it.k[q] = a.rank[g]; }
            seg.push_back(it);
        }
        auto less = [](const Item &p, const Item &q) {
            for (int w = 0; w < 3; w++) if (p.k[w] != q.k[w]) return p.k[w] < q.k[w];
            return false;
        };
        std::stable_sort(seg.begin(), seg.end(), less);
        *a.staged += size;
        u32 hd = 0;
        for (u32 x = 0; x <= size; x++) {
            if (x == size || (x > 0 && less(seg[x - 1], seg[x]))) { // the group [hd, x) is complete
                u32 c = x - hd;
                if (c == 1) a.singles[(*a.n_singles)++] = seg[hd].g;
                else { a.next[(*a.n_next)++] = ((u64)(start + hd) << 32) | c; if (c > *a.maxgroup) *a.maxgroup = c; }
                hd = x;
            }
            if (x < size) { a.sa[start + x] = seg[x].g; a.head[start + x] = start + hd; a.rank2[seg[x].g] = start + hd; }
        }
    }
}
#else
__global__ void __launch_bounds__(256) k_glist_build(GListBuildArgs a) {
    __shared__ u32 sm[33];
    __shared__ u32 s_base;
    const u32 i = blockIdx.x * 256u + threadIdx.x;
    u32 start = 0, size = 0;
    if (i < a.N && (i + 1 == a.N || (a.head[i + 1] & 0x7FFFFFFFu) == i + 1)) {
        start = a.head[i] & 0x7FFFFFFFu;
        size = i - start + 1;
    }
    const u32 f = size >= 2 ? 1u : 0u;
    u32 total;
    const u32 off = block_scan_excl(f, total, ScanSum(), sm);
    if (threadIdx.x == 0 && total) s_base = atomicAdd(a.count, total);
    __syncthreads();
    if (f) a.list[s_base + off] = ((u64)start << 32) | size;
}
static inline void launch_glist_build(Exec &ex, const GListBuildArgs &a) {
    if (a.N == 0) return;
    PROF_BEGIN(ex, "k_glist_build", 4.0 * a.N);
    k_glist_build<<<(a.N + 255) / 256, 256, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}

#define RG_GROUPS 32 // groups per CTA (4 per warp)
__global__ void __launch_bounds__(256) k_refine_g(RefineGArgs a) {
    __shared__ u64 s_new[RG_GROUPS * 16];
    __shared__ u32 s_sing[RG_GROUPS * 32];
    __shared__ u32 s_nnew, s_nsing, s_maxg, s_staged, s_base_new, s_base_sing;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, ltmask = (1u << lane) - 1u;
    if (tid == 0) { s_nnew = 0; s_nsing = 0; s_maxg = 1; s_staged = 0; }
    __syncthreads();
    const u32 g0 = blockIdx.x * RG_GROUPS;
    u32 widest = 0, staged = 0;
    // a warp's RG_GROUPS/8 groups are loaded side by side (every step of the chain list -> suffix ->
    // sequence -> ranks is a round trip to L2 or HBM), then ranked one after the other
    constexpr int GW = RG_GROUPS / 8;
    u32 gstart[GW], gsize[GW], gg[GW], gka[GW], gkb[GW], gkc[GW];
#pragma unroll
    for (int q = 0; q < GW; q++) {
        const u32 gi = g0 + warp + 8u * q;
        const u64 desc = (gi < a.nlist) ? a.list[gi] : 0ull;
        gstart[q] = (u32)(desc >> 32);
        gsize[q] = (u32)desc;
        staged += gsize[q];
    }
#pragma unroll
    for (int q = 0; q < GW; q++) gg[q] = (gsize[q] <= 32 && lane < gsize[q]) ? a.sa[gstart[q] + lane] : 0u;
    {
        u32 kq[GW], oq[GW], lq[GW];
#pragma unroll
        for (int q = 0; q < GW; q++) kq[q] = (gsize[q] <= 32 && lane < gsize[q]) ? seq_of(a.v, gg[q]) : 0u;
#pragma unroll
        for (int q = 0; q < GW; q++) {
            const bool m = gsize[q] <= 32 && lane < gsize[q];
            oq[q] = m ? LDG(a.v.seq_off + kq[q]) : 0u;
            lq[q] = m ? LDG(a.v.seq_off + kq[q] + 1) : 1u;
        }
#pragma unroll
        for (int q = 0; q < GW; q++) {
            gka[q] = 0xFFFFFFFFu; gkb[q] = 0; gkc[q] = 0;
            if (gsize[q] <= 32 && lane < gsize[q]) {
                const u32 off = oq[q], len = lq[q] - oq[q];
                u32 hh = a.h;
                if (hh >= len) hh %= len;
                u32 q1 = gg[q] - off + hh; if (q1 >= len) q1 -= len;
                gka[q] = LDG(a.rank + off + q1);
                if (a.nkeys == 3) {
                    u32 q2 = q1 + hh; if (q2 >= len) q2 -= len;
                    u32 q3 = q2 + hh; if (q3 >= len) q3 -= len;
                    gkb[q] = LDG(a.rank + off + q2);
                    gkc[q] = LDG(a.rank + off + q3);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < GW; q++) {
        const u32 start = gstart[q], size = gsize[q];
        if (size == 0) continue;
        if (size <= 32) { // one suffix per lane, everything in registers
            const bool member = lane < size;
            const u32 g = gg[q], ka = gka[q], kb = gkb[q], kc = gkc[q];
            unsigned rem = __ballot_sync(0xffffffffu, member);
            u32 placed = 0, mynew = 0, myhead = 0;
            bool alone = false;
            while (rem) {
                const u32 xa = (rem >> lane & 1u) ? ka : 0xFFFFFFFFu;
                const u32 ma = __reduce_min_sync(0xffffffffu, xa);
                const unsigned m1 = __ballot_sync(0xffffffffu, xa == ma) & rem;
                unsigned eq = m1;
                if (a.nkeys == 3 && (m1 & (m1 - 1))) { // several share the smallest first rank: look further
                    const u32 xb = (m1 >> lane & 1u) ? kb : 0xFFFFFFFFu;
                    const u32 mb = __reduce_min_sync(0xffffffffu, xb);
                    const unsigned m2 = __ballot_sync(0xffffffffu, xb == mb) & m1;
                    eq = m2;
                    if (m2 & (m2 - 1)) {
                        const u32 xc = (m2 >> lane & 1u) ? kc : 0xFFFFFFFFu;
                        const u32 mc = __reduce_min_sync(0xffffffffu, xc);
                        eq = __ballot_sync(0xffffffffu, xc == mc) & m2;
                    }
                }
                const u32 c = (u32)__popc(eq);
                if (eq >> lane & 1u) { mynew = placed + __popc(eq & ltmask); myhead = placed; alone = (c == 1); }
                if (c > 1 && (int)lane == __ffs((int)eq) - 1) {
                    s_new[atomicAdd(&s_nnew, 1u)] = ((u64)(start + placed) << 32) | c;
                    widest = c > widest ? c : widest;
                }
                placed += c;
                rem &= ~eq;
            }
            const unsigned sb = __ballot_sync(0xffffffffu, member && alone);
            u32 sbase = 0;
            if (lane == 0 && sb) sbase = atomicAdd(&s_nsing, (u32)__popc(sb));
            sbase = __shfl_sync(0xffffffffu, sbase, 0);
            if (member) {
                const u32 p = start + mynew, h2 = start + myhead;
                a.sa[p] = g;
                a.head[p] = h2;
                a.rank2[g] = h2;
                if (alone) s_sing[sbase + __popc(sb & ltmask)] = g;
            }
        } else { // several suffixes per lane; suffixes and their ranks parked in scratch at the group's places
            for (u32 t = lane; t < size; t += 32) {
                const u32 g = a.sa[start + t];
                const u32 k = seq_of(a.v, g);
                const u32 off = LDG(a.v.seq_off + k), len = LDG(a.v.seq_off + k + 1) - off;
                u32 hh = a.h;
                if (hh >= len) hh %= len;
                u32 q1 = g - off + hh; if (q1 >= len) q1 -= len;
                u32 q2 = q1 + hh; if (q2 >= len) q2 -= len;
                u32 q3 = q2 + hh; if (q3 >= len) q3 -= len;
                a.gs[start + t] = g;
                a.ka[start + t] = LDG(a.rank + off + q1);
                a.kb[start + t] = a.nkeys == 3 ? LDG(a.rank + off + q2) : 0u;
                a.kc[start + t] = a.nkeys == 3 ? LDG(a.rank + off + q3) : 0u;
            }
            __syncwarp();
            u32 placed = 0;
            while (placed < size) {
                u32 ma = 0xFFFFFFFFu, mb = 0xFFFFFFFFu, mc = 0xFFFFFFFFu;
                for (u32 t = lane; t < size; t += 32) { u32 w = a.ka[start + t]; ma = w < ma ? w : ma; } // placed ones are all ones
                ma = __reduce_min_sync(0xffffffffu, ma);
                for (u32 t = lane; t < size; t += 32)
                    if (a.ka[start + t] == ma) { u32 y = a.kb[start + t]; mb = y < mb ? y : mb; }
                mb = __reduce_min_sync(0xffffffffu, mb);
                for (u32 t = lane; t < size; t += 32)
                    if (a.ka[start + t] == ma && a.kb[start + t] == mb) { u32 y = a.kc[start + t]; mc = y < mc ? y : mc; }
                mc = __reduce_min_sync(0xffffffffu, mc);
                u32 cnt = 0;
                for (u32 t0 = 0; t0 < size; t0 += 32) { // how many share the smallest triple
                    const u32 t = t0 + lane;
                    const bool eq = t < size && a.ka[start + t] == ma && a.kb[start + t] == mb && a.kc[start + t] == mc;
                    cnt += __popc(__ballot_sync(0xffffffffu, eq));
                }
                u32 sbase = 0; // this subgroup's singleton (if it is one) or group goes straight to the global lists
                if (lane == 0) {
                    if (cnt == 1) sbase = atomicAdd(a.n_singles, 1u);
                    else { a.next[atomicAdd(a.n_next, 1u)] = ((u64)(start + placed) << 32) | cnt; widest = cnt > widest ? cnt : widest; }
                }
                sbase = __shfl_sync(0xffffffffu, sbase, 0);
                u32 seen = 0;
                for (u32 t0 = 0; t0 < size; t0 += 32) {
                    const u32 t = t0 + lane;
                    const bool eq = t < size && a.ka[start + t] == ma && a.kb[start + t] == mb && a.kc[start + t] == mc;
                    const unsigned b = __ballot_sync(0xffffffffu, eq);
                    if (eq) {
                        const u32 p = start + placed + seen + __popc(b & ltmask), h2 = start + placed;
                        const u32 g = a.gs[start + t];
                        a.sa[p] = g;
                        a.head[p] = h2;
                        a.rank2[g] = h2;
                        a.ka[start + t] = 0xFFFFFFFFu;
                        if (cnt == 1) a.singles[sbase] = g;
                    }
                    seen += __popc(b);
                }
                placed += cnt;
            }
        }
    }
    widest = __reduce_max_sync(0xffffffffu, widest);
    if (lane == 0) {
        if (widest > 1) atomicMax(&s_maxg, widest);
        if (staged) atomicAdd(&s_staged, staged);
    }
    __syncthreads();
    if (tid == 0) {
        s_base_new = s_nnew ? atomicAdd(a.n_next, s_nnew) : 0u;
        s_base_sing = s_nsing ? atomicAdd(a.n_singles, s_nsing) : 0u;
        atomicMax(a.maxgroup, s_maxg);
        atomicAdd(a.staged, s_staged);
    }
    __syncthreads();
    for (u32 i = tid; i < s_nnew; i += 256) a.next[s_base_new + i] = s_new[i];
    for (u32 i = tid; i < s_nsing; i += 256) a.singles[s_base_sing + i] = s_sing[i];
}
static inline void launch_refine_g(Exec &ex, const RefineGArgs &a) {
    if (a.nlist == 0) return;
    PROF_BEGIN(ex, "k_refine_g", 0.0);
    k_refine_g<<<(a.nlist + RG_GROUPS - 1) / RG_GROUPS, 256, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// ---- stage 1+2 fused, word sort: groups of the first sort ordered straight from the packed text -----
// After the first sort every group holds the rotations that share their first L0 letters -- in a set of
// related genomes mostly the m homologous copies of one place.  Prefix doubling would now shuffle RANKS
// through HBM round after round.  The packed, doubled text of a whole batch, however, is a few tens of MB
// and lives in the 126 MB L2, so the suffixes of a group can simply be compared letter word by letter word:
// every pair of suffixes of a group once (see ws_pairs below), WS_STEP words of 32 letters per round trip.
// A suffix's place in its group = the number of smaller ones; the most letters it shares with a smaller
// one IS its LCP (its predecessor is the smaller suffix it shares most with), so the LCP array falls out
// of the same pass (gencycsuffixtrees.c:500: never more than the shorter rotation -- pairs are followed
// to below the shortest sequence of the set; pairs that agree that far count as equal and their groups
// are left to the doubling rounds, as are groups of more than WS_BIG_CAP suffixes).
//   k_wsort     one warp per chunk: the groups that start inside WS_NOM consecutive places (<= 128 suffixes)
//   k_wsort_big one CTA per group that does not fit a warp's window (<= 1024 suffixes)
// HBM traffic: head 4 B per suffix in; sa 4 B in, sa + head + lcp 12 B out per suffix of a group.
// Cost: ~0.05 ns per pair on a B200.  Pairs grow with the square of the group, so this is the path for sets
// of a handful of genomes (groups of ~m) and the host falls back on rank doubling when the groups hold too
// many pairs (WS_PAIRS_PER_SUFFIX).  Another design was built and measured on the same data before this one
// and lost: splitting the groups 32 letters at a time in lock step, a warp per chunk (partition places by
// counting, then from ballots and popcounts): ~100 warp-synchronous passes per chunk with 2-3 % of the lanes
// busy in the tail -- 18-25 ms on 64 sets of 32 sequences, 2.8-4.2 ms on 160 Mammals-shaped sets where this
// takes 1.4.
#define WS_NOM 64    // SA places whose groups one warp takes (32: a third of the lanes had no pair to compare on the headline batch)
#define WS_CAP 128   // suffixes a warp can hold
#define WS_T 4
#define WS_WARPS 2   // warps (= chunks) per CTA of k_wsort: a CTA's registers and shared memory are free again only when its slowest
                     // warp is done, and most warps find no group and leave at once -- 8 warps a CTA: 3.95 ms on the headline batch, 4: 3.64, 2: 3.42
#define WSL_WARPS 8  // ... of k_wsort_list (every warp fetches groups until the list is done: no such tail)
#define WS_BIG_WARPS 8
#define WS_BIG_CAP (32 * WS_BIG_WARPS * WS_T)
#define WS_STEP 4    // words (of 32 letters) a pair compares per round trip
#define WS_DEPTH_CAP 32768u // letters a pair is followed; pairs that agree for longer are finished by the doubling rounds
#define WS_PAIRS_PER_SUFFIX 3.0 // the word sort is chosen when the groups of the first sort hold fewer pairs than this per suffix
#define WS_PAIRS_PER_SUFFIX_LARGE 10.0 // ... when the largest set of the batch has more than WS_LARGE_SET suffixes
#define WS_LARGE_SET (8u << 20)
#define LCP_UNKNOWN 0xFFFFFFFFu

struct WSortArgs {
    BatchView v; u32 *sa; u32 *head; u32 *lcp; u32 N;
    u32 lo, hi; // only the groups that start in [lo, hi) (both group borders): one rank's bucket, or 0, N
    u32 L0; u32 depth_cap; int masks;
    u64 *left; // groups still to be ordered (start : size), for the doubling rounds
    u64 *big;  // groups too long for a warp's window (start : size), for k_wsort_big
    u32 nbig;  // (k_wsort_big) entries of big
    u32 *res;  // [0] groups left, [1] suffixes in them, [2] fewest letters a left group shares, [3] largest left group,
               // [4] unused, [5] entries of big
    // the carried word sort (see "carry" below): groups as they stood after the first sort, which of them to take, and the
    // list of the groups sorted here whose order carries over to the groups one letter on
    const u32 *head_in;        // group borders are read here (nullptr: head)
    const unsigned char *flag; // by first place of a group (nullptr: every group is taken)
    u32 want;                  // ... the groups with flag[start] == want
    u32 *roots, *nroots;       // first places of the groups of <= CY_MAXG suffixes ordered here without a tie (nullptr: no list)
    u32 maxg;                  // CY_MAXG, or CY_BIGG: then the groups of CY_MAXG + 1 .. CY_BIGG suffixes go to a list of their own
    u32 *roots2, *nroots2;
    u32 big_min;               // (k_wsort_big) entries of big of up to big_min suffixes are k_wsort_words' (0: none)
    u32 list_cap;              // (k_wsort_list) groups of more suffixes than this go to big (0: WS_CAP, what a warp holds)
    u32 words_small;           // (k_wsort_list, ACGT only) groups of up to 32 suffixes word by word too (ws_words_warp)
};
HD const u32 *ws_head_in(const WSortArgs &a) { return a.head_in ? a.head_in : a.head; }
HD bool ws_taken(const WSortArgs &a, u32 start) { return !a.flag || a.flag[start] == a.want; }

HD u32 lexmask(u32 w) {
#if defined(__CUDA_ARCH__)
    return __brev(w);
#else
    u32 r = w;
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    r = ((r >> 2) & 0x33333333u) | ((r & 0x33333333u) << 2);
    r = ((r >> 4) & 0x0F0F0F0Fu) | ((r & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(r);
#endif
}
// (K, M): 2-bit letters and the "not ACGT" bits of one 32-letter word, both first letter on top; the
// fifth letter (code 4) is the largest and its 2-bit field is 0
HD bool ws_less(u64 ka, u32 ma, u64 kb, u32 mb) {
    const u32 dm = ma ^ mb;
    if (!dm) return ka < kb;
    const u64 dk = ka ^ kb;
    const int pm = CSA_CLZ(dm), pk = dk ? (CSA_CLZLL(dk) >> 1) : 32;
    if (pk < pm) return ka < kb;
    return (mb >> (31 - pm)) & 1u; // they part at a letter that is "other" in exactly one of them: that one is larger
}
HD u32 ws_common(u64 ka, u32 ma, u64 kb, u32 mb) { // letters the two words share (they differ)
    const u64 dk = ka ^ kb;
    const u32 dm = ma ^ mb;
    const u32 pk = dk ? (u32)(CSA_CLZLL(dk) >> 1) : 32u, pm = dm ? (u32)CSA_CLZ(dm) : 32u;
    return pk < pm ? pk : pm;
}

// ---- sharded run of ONE set: every rank sorts only the suffixes of its bucket ----------------------------------
// Buckets are ranges of the first 6 letters of the sort key (4096 prefixes): a histogram of the prefixes, the same
// on every rank, fixes the cut points and with them every bucket's place in the suffix array; a rank then keeps
// the (key, suffix) pairs of its prefixes -- in text order, the sort must stay stable -- and sorts those alone.
#define BK_BITS 12
#define BK_BINS (1 << BK_BITS)
struct BucketArgs {
    const u64 *keys64; const u32 *keys32; const u32 *vals; u32 N; int shift; // prefix = key >> shift
    u32 *hist;                       // k_bkhist: [BK_BINS]
    u32 plo, phi;                    // k_bkflag / k_bkscatter: this rank's prefixes [plo, phi)
    u32 *flag; const u32 *idx; u64 *out64; u32 *out32; u32 *outv;
};
HD u32 bk_prefix(const BucketArgs &a, long long i) { return a.keys32 ? (a.keys32[i] >> a.shift) : (u32)(a.keys64[i] >> a.shift); }
HD void bkflag_body(long long i, const BucketArgs &a) { const u32 p = bk_prefix(a, i); a.flag[i] = (p >= a.plo && p < a.phi) ? 1u : 0u; }
MAP_KERNEL(bkflag, BucketArgs, 8)
HD void bkscatter_body(long long i, const BucketArgs &a) {
    const u32 p = bk_prefix(a, i);
    if (p < a.plo || p >= a.phi) return;
    const u32 q = a.idx[i];
    if (a.keys32) a.out32[q] = a.keys32[i]; else a.out64[q] = a.keys64[i];
    a.outv[q] = a.vals[i];
}
MAP_KERNEL(bkscatter, BucketArgs, 12)
#ifdef CSA_EMU
HD void bkhist_body(long long i, const BucketArgs &a) { a.hist[bk_prefix(a, i)]++; }
MAP_KERNEL(bkhist, BucketArgs, 4)
#else
__global__ void __launch_bounds__(256) k_bkhist(BucketArgs a) { // counts in shared memory, one flush per CTA
    __shared__ u32 h[BK_BINS];
    for (u32 i = threadIdx.x; i < BK_BINS; i += 256) h[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.N; i += (long long)gridDim.x * 256) atomicAdd(&h[bk_prefix(a, i)], 1u);
    __syncthreads();
    for (u32 i = threadIdx.x; i < BK_BINS; i += 256) if (h[i]) atomicAdd(a.hist + i, h[i]);
}
static inline void launch_bkhist(Exec &ex, long long n, BucketArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_bkhist", (a.keys32 ? 4.0 : 8.0) * n);
    k_bkhist<<<148 * 4, 256, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// bucket borders for a job of nranks ranks: the first group border at or after r * N / nranks
struct BoundsArgs { const u32 *head; u32 N; u32 nranks; u32 *bounds; };
HD void bounds_map_body(long long r, const BoundsArgs &a) {
    u64 p = (u64)r * a.N / a.nranks;
    if (r == 0) p = 0;
    if ((u32)r >= a.nranks) p = a.N;
    if (p > 0 && p < a.N && (a.head[p] & 0x7FFFFFFFu) != (u32)p) { // inside a group: its end
        const u32 hs = a.head[p] & 0x7FFFFFFFu;
        u64 lo = p, step = 1;
        while (lo + step < a.N && (a.head[lo + step] & 0x7FFFFFFFu) == hs) { lo += step; step *= 2; }
        u64 hi = lo + step < a.N ? lo + step : a.N;
        while (hi - lo > 1) {
            const u64 mid = (lo + hi) >> 1;
            if ((a.head[mid] & 0x7FFFFFFFu) == hs) lo = mid; else hi = mid;
        }
        p = hi;
    }
    a.bounds[r] = (u32)p;
}
MAP_KERNEL(bounds_map, BoundsArgs, 4)
static inline void launch_bounds(Exec &ex, const BoundsArgs &a) { launch_bounds_map(ex, (long long)a.nranks + 1, a); }

// ---- the carried word sort: one compare per column of near-identical genomes, not one per place ----------------
// Two rotations that share l > L0 letters still share l - 1 one place on, in the same order.  So when every suffix of
// a group G' has its predecessor in the text (one letter back round its sequence) in ONE group G of the first sort,
// G' is G's order, moved one place, restricted to the suffixes whose L0 letters still agree -- and the LCPs are G's
// less one (Kasai's carry-over, applied to whole groups).  Only the groups that are NOT of that kind ("roots": a
// predecessor differs, i.e. the column follows a place where the genomes differ) are ordered by comparing letters
// (k_wsort on the flagged groups); one warp per root then walks the text, a lane per suffix, and writes the groups
// that follow from it, step by step, until every lane's run has ended:
//   k_cygrp    grp[suffix] = first place of its group (unset: alone in its group, or the group holds more than a
//              warp); snapshot of the heads (the walk reads borders while other warps rewrite heads)
//   k_cyroots  flag[first place] = 1 for the groups that must be ordered afresh: a predecessor in another group or
//              in none, or its first suffix's number a multiple of CY_CUT (cuts the walks,
//              so that ten thousand of them run side by side whatever the genomes share)
//   k_cylist + k_wsort_list   (want = 1) orders the roots by letters -- word by word, a warp a group (ws_words_warp), when the
//              batch is ACGT only, else pair by pair -- and lists those a walk can start from (no two suffixes still equal)
//   k_cywalk   the walks; flag = 2 on every group written
//   k_cylist + k_wsort_list   (want = 0) whatever no walk reached (descendants of roots with ties): as before, by letters
// Sets of hundreds of near-identical sequences (a.maxg == CY_BIGG): groups of up to 256 suffixes are walked too, a CTA a
// root (k_cywalk_cta, a second list of roots); roots of more than 32 suffixes are ordered by k_wsort_words, a CTA a group.
// Results are those of the word sort alone, place by place (tests: forced on every golden set; full-size agreement).
#define CY_MAXG 32u   // suffixes a warp's walk carries
#define CY_BIGG 256u  // ... a CTA's walk (k_cywalk_cta: sets of hundreds of near-identical sequences)
#define CY_CUT 1024u
#define CY_UNSET 0xFFFFFFFFu
struct CarryArgs {
    BatchView v; const u32 *sa; const u32 *head; u32 *head2; u32 *grp; unsigned char *flag; u32 lo, hi;
    int pack;  // 1 (batches of < 2^27 suffixes): a grp entry = first place << 5 | size - 1, the walks need no second look at the heads
    u32 *gval; // nullptr: grp[suffix] written at once; else the value by place -- one set of tens of millions of suffixes:
               // 4-byte stores all over a 320 MB array cost 2.9 ms, dealt by the top 8 bits of the suffix first (one
               // radix pass) and stored then (k_cyscatter), every stretch of the array is filled while it sits in L2
    u32 maxg;  // groups of up to maxg suffixes are walked: CY_MAXG, or CY_BIGG (then pack == 0)
    const u32 *startbits; // bit s: suffix s is the first of its sequence
};
HD bool cy_single(const u32 *h, u32 x, u32 hs, u32 hi) { return hs == x && (x + 1 >= hi || (h[x + 1] & 0x7FFFFFFFu) != hs); }
HD bool cy_big(const u32 *h, u32 hs, u32 hi, u32 maxg) { return (u64)hs + maxg < hi && (h[hs + maxg] & 0x7FFFFFFFu) == hs; }
HD void cygrp_body(long long i, const CarryArgs &a) {
    const u32 x = a.lo + (u32)i, hs = a.head[x] & 0x7FFFFFFFu;
    a.head2[x] = hs;
    const bool big = cy_big(a.head, hs, a.hi, a.maxg), single = cy_single(a.head, x, hs, a.hi);
    u32 val = (big || single) ? CY_UNSET : hs;
    if (a.pack && val != CY_UNSET) {
        u32 e = x + 1;
        while (e < a.hi && (a.head[e] & 0x7FFFFFFFu) == hs) e++; // (at most 31 steps: the group holds no more than 32; pack only with maxg == CY_MAXG)
        val = (hs << 5) | (e - hs - 1u);
    }
    if (a.gval) a.gval[i] = val; else a.grp[a.sa[x]] = val;
    a.flag[x] = (hs != x || single) ? 3 : big ? 1 : 0; // 3: no first place of a group of two or more (k_cylist reads the flags alone there)
}
#ifdef CSA_EMU
MAP_KERNEL(cygrp, CarryArgs, 17)
#else
// the same; a group's end from the borders of the warp's 32 places and the 32 behind them (two ballots), no walk along the heads
__global__ void __launch_bounds__(256) k_cygrp(long long n, CarryArgs a) {
    const u32 lane = threadIdx.x & 31u;
#pragma unroll 1
    for (int j = 0; j < MAP_ITEMS; j++) { // (MAP_ITEMS stretches of 256 places a CTA: see MAP_KERNEL_N)
    const long long i = (long long)blockIdx.x * (256 * MAP_ITEMS) + 256 * j + threadIdx.x;
    if (i - lane >= n) break; // (the whole warp at once: the ballots below are the warp's)
    const bool in = i < n;
    const u64 x64 = (u64)a.lo + (u64)(in ? i : 0), y64 = x64 + 32u;
    const u32 x = (u32)x64;
    const u32 hs = in ? a.head[x] & 0x7FFFFFFFu : 0u;
    const u32 hy = (in && y64 < a.hi) ? a.head[y64] & 0x7FFFFFFFu : 0u;
    // border bits: place p starts a group (the end of the range counts as one)
    const u64 b = (u64)__ballot_sync(0xffffffffu, !in || hs == x) | ((u64)__ballot_sync(0xffffffffu, !in || y64 >= a.hi || hy == (u32)y64) << 32);
    if (!in) continue;
    a.head2[x] = hs;
    const u64 above = b & (~0ull << (lane + 1u));
    const u32 e = above ? x - lane + (u32)(__ffsll((long long)above) - 1) : x - lane + 64u; // first border behind x (at most 63 places on)
    const bool big = e - hs > CY_MAXG, single = e - hs == 1u;
    u32 val = (big || single) ? CY_UNSET : hs;
    if (a.pack && val != CY_UNSET) val = (hs << 5) | (e - hs - 1u);
    if (a.gval) a.gval[i] = val; else a.grp[a.sa[x]] = val;
    a.flag[x] = (hs != x || single) ? 3 : big ? 1 : 0; // 3: no first place of a group of two or more (k_cylist reads the flags alone there)
    }
}
HD void cygrp_any_body(long long i, const CarryArgs &a) { cygrp_body(i, a); }
MAP_KERNEL(cygrp_any, CarryArgs, 17) // (groups of up to CY_BIGG: the group's end is not among the 64 places of two ballots)
static inline void launch_cygrp(Exec &ex, long long n, CarryArgs a) {
    if (n <= 0) return;
    if (a.maxg != CY_MAXG) { launch_cygrp_any(ex, n, a); return; }
    PROF_BEGIN(ex, "k_cygrp", 17.0 * n);
    k_cygrp<<<(unsigned)((n + 256 * MAP_ITEMS - 1) / (256 * MAP_ITEMS)), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif
struct CyScatterArgs { const u32 *suffix; const u32 *val; u32 *grp; };
HD void cyscatter_body(long long i, const CyScatterArgs &a) { a.grp[a.suffix[i]] = a.val[i]; }
MAP_KERNEL_N(cyscatter, CyScatterArgs, 12)
// the group (first place) of the suffix one letter back round its sequence; *cut: the suffix's number is a multiple of CY_CUT.
// Which suffixes are the first of their sequence (their letter back is the sequence's last) is a bit set of N bits, L2-resident
// (k_cyseqbits): all others need no look at their sequence -- a gather from an N-sized table for every suffix of the batch
// (k_cyroots 1.17 -> 0.77 ms on 192 sets of 32).
HD u32 cy_parent(const CarryArgs &a, u32 s, bool *cut) {
    *cut = (s & (CY_CUT - 1u)) == 0u;
    if (a.v.M <= 64u) { // a few long sequences: their starts sit in L1, a search there is cheaper than the bit
        const u32 k = seq_of_few(a.v, s);
        return a.grp[s != LDG(a.v.seq_off + k) ? s - 1u : LDG(a.v.seq_off + k + 1) - 1u];
    }
    if (!(LDG(a.startbits + (s >> 5)) >> (s & 31u) & 1u)) return a.grp[s - 1u];
    return a.grp[LDG(a.v.seq_off + seq_of(a.v, s) + 1) - 1u];
}
struct CySeqBitsArgs { BatchView v; u32 *startbits; };
HD void cyseqbits_body(long long k, const CySeqBitsArgs &a) { const u32 s = LDG(a.v.seq_off + k); ATOMIC_OR(a.startbits + (s >> 5), 1u << (s & 31u)); }
MAP_KERNEL(cyseqbits, CySeqBitsArgs, 8)
#ifdef CSA_EMU
HD void cyroots_body(long long i, const CarryArgs &a) {
    const u32 x = a.lo + (u32)i, hs = a.head2[x];
    if (cy_single(a.head2, x, hs, a.hi) || cy_big(a.head2, hs, a.hi, a.maxg)) return;
    bool cut, cut0;
    const u32 par = cy_parent(a, a.sa[x], &cut);
    bool root = (cut && x == hs) || par == CY_UNSET; // (the cut: by the group's first suffix -- the smallest, the first sort is stable)
    if (!root && x != hs) root = par != cy_parent(a, a.sa[hs], &cut0);
    if (root) a.flag[hs] = 1;
}
MAP_KERNEL(cyroots, CarryArgs, 20)
#else
// the same; "all predecessors in one group" as "every suffix's predecessor in the group of its neighbour's": the neighbour's
// comes by a shuffle, only a warp's first lane fetches it
__global__ void __launch_bounds__(256) k_cyroots(long long n, CarryArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 lane = threadIdx.x & 31u;
    const bool in = i < n;
    const u32 x = a.lo + (u32)(in ? i : 0), hs = in ? a.head2[x] : 0xFFFFFFFEu;
    const bool take = in && !cy_single(a.head2, x, hs, a.hi) && !cy_big(a.head2, hs, a.hi, a.maxg);
    bool cut = false, cut0;
    const u32 par = take ? cy_parent(a, a.sa[x], &cut) : CY_UNSET;
    u32 prev = __shfl_up_sync(0xffffffffu, par, 1);
    if (take && x != hs && lane == 0) prev = cy_parent(a, a.sa[x - 1], &cut0); // (x - 1 is of the same group: not alone, not big)
    const bool root = take && ((cut && x == hs) || par == CY_UNSET || (x != hs && par != prev));
    if (root) a.flag[hs] = 1;
}
static inline void launch_cyroots(Exec &ex, long long n, CarryArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_cyroots", 20.0 * n);
    k_cyroots<<<(unsigned)((n + 255) / 256), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif
// the groups of two or more with flag == want, as a list (first place : size) -- the few that are ordered by letters
struct CyListArgs { const u32 *head2; const unsigned char *flag; u32 want; u32 lo, hi; u64 *list; u32 *count; };
HD void cylist_body(long long i, const CyListArgs &a) {
    const u32 x = a.lo + (u32)i;
    if (a.flag[x] != a.want) return; // (first places of groups of two or more only: k_cygrp marks all others)
    if (a.head2[x] != x || cy_single(a.head2, x, x, a.hi)) return;
    u64 e = (u64)x + 2;
    while (e < a.hi && e < (u64)x + CY_MAXG && a.head2[e] == x) e++;
    if (e < a.hi && a.head2[e] == x) { // longer than a warp: gallop, then bisect
        u64 lo = e, step = CY_MAXG;
        while (lo + step < a.hi && a.head2[lo + step] == x) { lo += step; step *= 2; }
        u64 hi = lo + step < a.hi ? lo + step : a.hi;
        while (hi - lo > 1) { const u64 mid = (lo + hi) >> 1; if (a.head2[mid] == x) lo = mid; else hi = mid; }
        e = hi;
    }
    a.list[ATOMIC_ADD(a.count, 1u)] = ((u64)x << 32) | (u32)(e - x);
}
#ifdef CSA_EMU
MAP_KERNEL(cylist, CyListArgs, 1)
#else
// the same, 16 places a thread: the flags are read as one 16-byte word, and almost none of them is wanted (a thread a place
// is 400 000 CTAs that read one byte each and leave: the launch rate, not the 100 MB, set the 0.26 ms)
__global__ void __launch_bounds__(256) k_cylist(long long nthreads, CyListArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const u64 x0 = (u64)(a.lo & ~15u) + 16ull * (u64)t;
    const uint4 f = *reinterpret_cast<const uint4 *>(a.flag + x0); // (the flags of a batch: a 256-byte aligned buffer of its own, padded)
    const u32 w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        // bytes equal to want: x ^ want*0x01010101 has a zero byte there
        const u32 v = w[q] ^ (a.want * 0x01010101u);
        if (((v - 0x01010101u) & ~v & 0x80808080u) == 0u) continue; // no zero byte
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const u64 x = x0 + 4u * q + b;
            if (((v >> (8 * b)) & 0xFFu) == 0u && x >= a.lo && x < a.hi) cylist_body((long long)(x - a.lo), a);
        }
    }
}
static inline void launch_cylist(Exec &ex, long long n, CyListArgs a) {
    if (n <= 0) return;
    const long long nthreads = ((long long)a.hi - (long long)(a.lo & ~15u) + 15) / 16;
    PROF_BEGIN(ex, "k_cylist", 1.0 * n);
    k_cylist<<<(unsigned)((nthreads + 255) / 256), 256, 0, ex.stream>>>(nthreads, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

struct CyWalkArgs {
    BatchView v; u32 *sa; u32 *head; u32 *lcp; const u32 *head2; const u32 *grp; int pack; unsigned char *flag;
    u32 lo, hi, L0; const u32 *roots; const u32 *nroots; u32 *next; // next: the walks' work counter (zeroed)
    const u32 *roots2; const u32 *nroots2; u32 *next2; // roots of more than CY_MAXG suffixes (k_cywalk_cta); nullptr: there are none
};
#ifdef CSA_EMU
// one root, lane by lane as the warp does it
static inline void emu_cywalk_root(const CyWalkArgs &a, u32 hs) {
    u32 g = 1;
    while (hs + g < a.hi && a.head2[hs + g] == hs) g++;
    struct Lane { u32 st, len, off, lcp; bool on; };
    std::vector<Lane> ln(g);
    for (u32 r = 0; r < g; r++) {
        const u32 s = a.sa[hs + r], k = seq_of(a.v, s), st = a.v.seq_off[k];
        ln[r] = Lane{st, a.v.seq_off[k + 1] - st, s - st, r ? a.lcp[hs + r] : 0u, true};
    }
    for (u32 j = 1;; j++) {
        bool any = false;
        for (u32 r = 0; r < g; r++)
            if (ln[r].on) { ln[r].off = ln[r].off + 1 == ln[r].len ? 0 : ln[r].off + 1; any = true; }
        if (!any) return;
        for (u32 rs = 0; rs < g;) {
            if (!ln[rs].on) { rs++; continue; }
            u32 re = rs + 1;
            while (re < g && ln[re].on && ln[re].lcp >= a.L0 + j) re++;
            const u32 gv = a.grp[ln[rs].st + ln[rs].off], n = re - rs;
            const u32 G = (a.pack && gv != CY_UNSET) ? gv >> 5 : gv;
            bool ok = gv != CY_UNSET && n >= 2;
            if (ok) ok = a.pack ? n == (gv & 31u) + 1u : ((u64)G + n >= a.hi || a.head2[G + n] != G);
            if (ok) { // the cut: by the smallest suffix of the group
                u32 first = rs;
                for (u32 r = rs + 1; r < re; r++) if (ln[r].st + ln[r].off < ln[first].st + ln[first].off) first = r;
                if (((ln[first].st + ln[first].off) & (CY_CUT - 1u)) == 0u) ok = false;
            }
            if (ok) {
                for (u32 r = rs; r < re; r++) {
                    const u32 place = G + (r - rs);
                    a.sa[place] = ln[r].st + ln[r].off;
                    a.head[place] = place;
                    if (r > rs) a.lcp[place] = ln[r].lcp - j;
                }
                a.flag[G] = 2;
            } else for (u32 r = rs; r < re; r++) ln[r].on = false;
            rs = re;
        }
    }
}
static inline void launch_cywalk(Exec &, const CyWalkArgs &a) {
    for (u32 i = 0; i < *a.nroots; i++) emu_cywalk_root(a, a.roots[i]);
}
static inline void launch_cywalk_cta(Exec &, const CyWalkArgs &a) {
    for (u32 i = 0; i < *a.nroots2; i++) emu_cywalk_root(a, a.roots2[i]);
}
#else
#define CY_WARPS 8
__global__ void __launch_bounds__(CY_WARPS * 32) k_cywalk(CyWalkArgs a) {
    const u32 lane = threadIdx.x & 31u, nroots = *a.nroots;
    for (;;) {
        u32 i = 0;
        if (lane == 0) i = atomicAdd(a.next, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nroots) return;
        const u32 hs = a.roots[i];
        const bool mine = (u64)hs + lane < a.hi && a.head2[hs + lane] == hs; // (a root holds at most 32 suffixes)
        u32 st = 0, len = 1, off = 0, l = 0;
        if (mine) {
            const u32 s = a.sa[hs + lane], k = seq_of_few(a.v, s);
            st = LDG(a.v.seq_off + k); len = LDG(a.v.seq_off + k + 1) - st; off = s - st;
            l = lane ? a.lcp[hs + lane] : 0u;
        }
        bool on = mine;
        off = off + 1u == len ? 0u : off + 1u;
        u32 gnext = on ? LDG(a.grp + st + off) : CY_UNSET; // (the entry of the next step is fetched a step ahead)
        for (u32 j = 1;; j++) {
            const u32 onmask = __ballot_sync(0xffffffffu, on);
            if (!onmask) break;
            const u32 gv = gnext, pos = st + off;
            off = off + 1u == len ? 0u : off + 1u;
            gnext = on ? LDG(a.grp + st + off) : CY_UNSET;
            const u32 G = (a.pack && gv != CY_UNSET) ? gv >> 5 : gv;
            // runs: lane r goes on with lane r - 1 when both are in the walk and still share L0 letters after j steps
            const bool cont = on && lane && ((onmask >> (lane - 1u)) & 1u) && l >= a.L0 + j;
            const u32 startmask = __ballot_sync(0xffffffffu, on && !cont);
            bool ok = false;
            u32 rs = 0, n = 0, runmask = 0;
            bool cutme = false;
            if (on) {
                rs = 31u - (u32)__clz((int)(startmask & (0xFFFFFFFFu >> (31u - lane))));
                const u32 ends = (startmask | ~onmask) & (lane < 31u ? 0xFFFFFFFFu << (lane + 1u) : 0u);
                const u32 re = ends ? (u32)__ffs((int)ends) - 1u : 32u;
                n = re - rs;
                runmask = (re < 32u ? (1u << re) - 1u : 0xFFFFFFFFu) & (0xFFFFFFFFu << rs);
                // the cut: by the smallest suffix of the run (the lanes of a run hold the same mask)
                cutme = __reduce_min_sync(runmask, pos) == pos && (pos & (CY_CUT - 1u)) == 0u;
            }
            const u32 cutmask = __ballot_sync(0xffffffffu, cutme);
            if (on) {
                ok = gv != CY_UNSET && n >= 2u && !(cutmask & runmask);
                if (ok) ok = a.pack ? n == (gv & 31u) + 1u : ((u64)G + n >= a.hi || LDG(a.head2 + G + n) != G); // the group holds nobody else
            }
            if (ok) {
                const u32 place = G + (lane - rs);
                a.sa[place] = pos;
                a.head[place] = place;
                if (lane > rs) a.lcp[place] = l - j; else a.flag[G] = 2;
            }
            on = ok;
        }
    }
}
static inline void launch_cywalk(Exec &ex, const CyWalkArgs &a) {
    PROF_BEGIN(ex, "k_cywalk", 0.0);
    k_cywalk<<<148 * 8, CY_WARPS * 32, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
// the same walk for a root of up to CY_BIGG suffixes: one CTA, a thread per suffix.  What the warp's walk reads off two
// ballots -- where the runs start and end -- is read off the warps' ballot words in shared memory; the cut (a run whose
// smallest suffix stands at a multiple of CY_CUT ends the walk) is settled only in the steps that have such a suffix at
// all: the candidates leave the smallest of them at their run's first place, whoever of the run is smaller still vetoes.
#define CYB_WARPS (CY_BIGG / 32u)
__global__ void __launch_bounds__(CY_BIGG) k_cywalk_cta(CyWalkArgs a) {
    __shared__ u32 s_on[CYB_WARPS], s_start[CYB_WARPS], s_root, s_cut[CY_BIGG], s_veto[CY_BIGG];
    const u32 tid = threadIdx.x, lane = tid & 31u, w = tid >> 5, nroots = *a.nroots2;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_root = atomicAdd(a.next2, 1u);
        __syncthreads();
        const u32 i = s_root;
        if (i >= nroots) return;
        const u32 hs = a.roots2[i];
        const bool mine = (u64)hs + tid < a.hi && a.head2[hs + tid] == hs; // (such a root holds at most CY_BIGG suffixes)
        u32 st = 0, len = 1, off = 0, l = 0;
        if (mine) {
            const u32 s = a.sa[hs + tid], k = seq_of_few(a.v, s);
            st = LDG(a.v.seq_off + k); len = LDG(a.v.seq_off + k + 1) - st; off = s - st;
            l = tid ? a.lcp[hs + tid] : 0u;
        }
        bool on = mine;
        off = off + 1u == len ? 0u : off + 1u;
        u32 gnext = on ? LDG(a.grp + st + off) : CY_UNSET;
        for (u32 j = 1;; j++) {
            const u32 onmask = __ballot_sync(0xffffffffu, on);
            if (lane == 0) s_on[w] = onmask;
            __syncthreads();
            s_cut[tid] = 0xFFFFFFFFu; s_veto[tid] = 0u; // (last read before this barrier, next written behind the two that follow)
            u32 anyon = 0;
#pragma unroll
            for (u32 q = 0; q < CYB_WARPS; q++) anyon |= s_on[q];
            if (!anyon) break;
            const u32 gv = gnext, pos = st + off;
            off = off + 1u == len ? 0u : off + 1u;
            gnext = on ? LDG(a.grp + st + off) : CY_UNSET;
            const u32 G = gv; // (pack == 0)
            const bool prev_on = lane ? ((onmask >> (lane - 1u)) & 1u) != 0u : (w != 0u && (s_on[w - 1u] >> 31) != 0u);
            const bool cont = on && prev_on && l >= a.L0 + j;
            const u32 startmask = __ballot_sync(0xffffffffu, on && !cont);
            if (lane == 0) s_start[w] = startmask;
            __syncthreads();
            u32 rs = 0, n = 0;
            if (on) { // first and last place of my run: the start at or before me, the first start or gap behind me
                u32 m = startmask & (0xFFFFFFFFu >> (31u - lane)), ww = w;
                while (!m) m = s_start[--ww]; // (a run has a start: a place in the walk whose neighbour before is not goes on with nobody)
                rs = ww * 32u + 31u - (u32)__clz((int)m);
                u32 e = (startmask | ~onmask) & (lane < 31u ? 0xFFFFFFFFu << (lane + 1u) : 0u);
                ww = w;
                while (!e && ++ww < CYB_WARPS) e = s_start[ww] | ~s_on[ww];
                const u32 re = e ? ww * 32u + (u32)__ffs((int)e) - 1u : CY_BIGG;
                n = re - rs;
            }
            const bool cand = on && (pos & (CY_CUT - 1u)) == 0u;
            bool cutrun = false;
            if (__syncthreads_or(cand)) {
                if (cand) atomicMin(&s_cut[rs], pos);
                __syncthreads();
                if (on && pos < s_cut[rs]) s_veto[rs] = 1u; // (s_cut unset = 0xFFFFFFFF: a veto nobody reads)
                __syncthreads();
                cutrun = on && s_cut[rs] != 0xFFFFFFFFu && !s_veto[rs];
            }
            bool ok = on && gv != CY_UNSET && n >= 2u && !cutrun;
            if (ok) ok = (u64)G + n >= a.hi || LDG(a.head2 + G + n) != G; // the group holds nobody else
            if (ok) {
                const u32 place = G + (tid - rs);
                a.sa[place] = pos;
                a.head[place] = place;
                if (tid > rs) a.lcp[place] = l - j; else a.flag[G] = 2;
            }
            on = ok;
        }
    }
}
static inline void launch_cywalk_cta(Exec &ex, const CyWalkArgs &a) {
    PROF_BEGIN(ex, "k_cywalk_cta", 0.0);
    k_cywalk_cta<<<148 * 8, CY_BIGG, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
#endif

#ifdef CSA_EMU
// one group [p, e) by the letters [L0, Lend): stable; what is still together goes to the list
static inline void emu_wsort_group(const WSortArgs &a, u32 p, u32 e, u32 Lend, bool with_root = false) {
    const BatchView &v = a.v;
    struct It { u32 g; std::vector<unsigned char> s; };
    std::vector<It> items;
    for (u32 x = p; x < e; x++) {
        It it{a.sa[x], {}};
        for (u32 h = a.L0; h < Lend; h++) it.s.push_back(v.code[cyc_add(v, it.g, h)]);
        items.push_back(std::move(it));
    }
    std::stable_sort(items.begin(), items.end(), [](const It &x, const It &y) { return x.s < y.s; });
    u32 hd = p;
    bool ties = false;
    for (u32 x = p; x <= e; x++) {
        const bool brk = x == e || (x > p && items[x - p - 1].s != items[x - p].s);
        if (brk) {
            if (x - hd >= 2) { // still together after Lend letters
                ties = true;
                a.left[a.res[0]++] = ((u64)hd << 32) | (x - hd);
                a.res[1] += x - hd;
                if (Lend < a.res[2]) a.res[2] = Lend;
                if (x - hd > a.res[3]) a.res[3] = x - hd;
            }
            if (x < e) {
                const auto &s1 = items[x - p - 1].s, &s2 = items[x - p].s;
                u32 c = 0;
                while (s1[c] == s2[c]) c++;
                a.lcp[x] = a.L0 + c;
            }
            hd = x;
        }
        if (x < e) { a.sa[x] = items[x - p].g; a.head[x] = hd; }
    }
    if (with_root && a.roots && !ties && e - p <= CY_MAXG) a.roots[(*a.nroots)++] = p;
    else if (with_root && a.roots && !ties && e - p <= a.maxg) a.roots2[(*a.nroots2)++] = p;
}
static inline void emu_wsort_leave(const WSortArgs &a, u32 p, u32 e) {
    a.left[a.res[0]++] = ((u64)p << 32) | (e - p);
    a.res[1] += e - p;
    if (a.L0 < a.res[2]) a.res[2] = a.L0;
    if (e - p > a.res[3]) a.res[3] = e - p;
}
static inline u32 emu_wsort_lend(const WSortArgs &a, u32 nmin) {
    const u32 Lmax = nmin < a.depth_cap ? nmin : a.depth_cap;
    u32 Lend = a.L0;
    while (Lend + 32 <= Lmax) Lend += 32;
    return Lend;
}
static inline void launch_wsort(Exec &, const WSortArgs &a) {
    const BatchView &v = a.v;
    const u32 N = a.N;
    (void)N; // (heads beyond hi may belong to another rank's bucket and not be there yet: hi itself is a border)
    const u32 *hin = ws_head_in(a);
    auto border = [&](u64 p) { return p < a.hi ? (hin[p] & 0x7FFFFFFFu) == (u32)p : p == a.hi; };
    for (u64 r0 = a.lo & ~31u; r0 < a.hi; r0 += WS_NOM) {
        // groups that start in [r0, r0+WS_NOM) and end at or before place r0 + WS_CAP - 1
        std::vector<std::pair<u32, u32>> groups;
        u32 nmin = 0xFFFFFFFFu;
        for (u64 p = r0; p < r0 + WS_NOM && p < a.hi; p++) {
            if (p < a.lo || !border(p)) continue;
            u64 e = p + 1;
            while (!border(e)) e++;
            if (e - r0 > WS_CAP - 1) { // runs past the warp's window
                if (e - p >= 2 && ws_taken(a, (u32)p)) a.big[a.res[5]++] = ((u64)p << 32) | (u32)(e - p);
                break;
            }
            if (e - p >= 2 && ws_taken(a, (u32)p)) {
                groups.push_back({(u32)p, (u32)e});
                for (u64 x = p; x < e; x++) {
                    u32 nm = v.set_nmin[v.seq_set[v.seqof[a.sa[x]]]];
                    if (nm < nmin) nmin = nm;
                }
            }
        }
        for (auto &gr : groups) emu_wsort_group(a, gr.first, gr.second, emu_wsort_lend(a, nmin), true);
    }
}
// the groups of a list (k_cylist) instead of every group of [lo, hi)
static inline void launch_wsort_list(Exec &, const WSortArgs &a, const u64 *list, const u32 *count) {
    for (u32 i = 0; i < *count; i++) {
        const u32 p = (u32)(list[i] >> 32), e = p + (u32)list[i];
        if (e - p > (a.list_cap ? a.list_cap : (u32)WS_CAP)) { a.big[a.res[5]++] = list[i]; continue; }
        u32 nmin = 0xFFFFFFFFu;
        for (u32 x = p; x < e; x++) { const u32 nm = a.v.set_nmin[a.v.seq_set[a.v.seqof[a.sa[x]]]]; if (nm < nmin) nmin = nm; }
        emu_wsort_group(a, p, e, emu_wsort_lend(a, nmin), true);
    }
}
static inline void launch_wsort_big(Exec &, const WSortArgs &a) {
    for (u32 i = 0; i < a.nbig; i++) {
        const u32 p = (u32)(a.big[i] >> 32), e = p + (u32)a.big[i];
        if (e - p > WS_BIG_CAP) { emu_wsort_leave(a, p, e); continue; }
        const u32 nmin = a.v.set_nmin[a.v.seq_set[a.v.seqof[a.sa[p]]]];
        emu_wsort_group(a, p, e, emu_wsort_lend(a, nmin), a.roots != nullptr);
    }
}
#else
HD u32 ws_next_border(u64 b0, u64 b1, u32 t) { // first border place after t (place WS_CAP always counts as one)
    if (t < 63) {
        const u64 m = b0 & (~0ull << (t + 1));
        if (m) return (u32)(__ffsll((long long)m) - 1);
        return b1 ? 64u + (u32)(__ffsll((long long)b1) - 1) : (u32)WS_CAP;
    }
    if (t == 63) return b1 ? 64u + (u32)(__ffsll((long long)b1) - 1) : (u32)WS_CAP;
    const u32 tt = t - 64;
    const u64 m = tt < 63 ? (b1 & (~0ull << (tt + 1))) : 0ull;
    return m ? 64u + (u32)(__ffsll((long long)m) - 1) : (u32)WS_CAP;
}

// a team = 1 warp (k_wsort) or the whole CTA (k_wsort_big); place t of its chunk belongs to thread t % threads
template <int WARPS> struct WsTeam {
    static __device__ __forceinline__ void sync() { if (WARPS == 1) __syncwarp(); else __syncthreads(); }
};
template <int WARPS> struct WsSmem {
    static constexpr int CAP = 32 * WARPS * WS_T;
    u64 x[CAP];                 // where the suffix at place t starts in the doubled text
    u32 g[CAP];                 // the suffix
    u32 ct[CAP];                // low half: suffixes of its group that are smaller; high half: equal ones that stood before it
    u32 best[CAP];              // most letters shared with a smaller one = LCP with its predecessor
    u32 clsz[CAP];              // (by new place) suffixes that are still together there
    u32 pre[CAP + 1];           // pairs that the places before t bring (running sum); [CAP] = all pairs of the chunk
    unsigned short seg[CAP];    // first place of its group
    unsigned short end[CAP];    // place after the last of its group
};

// Every pair of suffixes of a group is compared once, word by word from letter L0 on (suffix i of a group
// of g brings the pairs with i+1 .. i+(g-1)/2 round the group).  The pairs of the whole chunk are numbered
// through a running sum and dealt out to the threads in turn, so every lane has work whichever places hold
// the groups; a thread walks its pairs on its own: lanes are in different pairs at different depths and
// never wait for one another -- no ballot, no barrier until all pairs are done.  What a pair tells: the smaller suffix counts towards the larger one's place, and the letters they
// share bound the larger one's LCP from below (its predecessor is the smaller suffix it shares most
// with).  Pairs that agree up to the depth limit count as equal (place by old order; the doubling rounds
// finish them).  On entry: x, g, seg, end set and ct = best = clsz = 0 for every place with act[j].
template <bool MASKS, int WARPS, bool ROOTS = false>
__device__ __forceinline__ void ws_pairs(const WSortArgs &a, WsSmem<WARPS> &s, const u32 tid, const u32 base,
                                         const bool (&act)[WS_T], const u32 Lmax) {
    typedef WsTeam<WARPS> Team;
    constexpr u32 TT = 32u * WARPS;
    const u32 L0 = a.L0;
    const u32 Lend = Lmax >= L0 ? L0 + ((Lmax - L0) & ~31u) : L0; // pairs that agree on [L0, Lend) count as equal
    constexpr u32 CAP = (u32)WsSmem<WARPS>::CAP, PER = CAP / 32u;
#pragma unroll
    for (int j = 0; j < WS_T; j++) { // pairs brought by every place
        const u32 t = tid + TT * j;
        u32 nd = 0;
        if (act[j]) {
            const u32 hs = s.seg[t], g = (u32)s.end[t] - hs, i = t - hs;
            nd = (g - 1u) / 2u + (((g & 1u) == 0u && i < g / 2u) ? 1u : 0u);
        }
        s.pre[t] = nd;
    }
    Team::sync();
    if (tid < 32u) { // running sum over the places (first warp of the team)
        u32 sum = 0;
#pragma unroll
        for (u32 q = 0; q < PER; q++) sum += s.pre[tid * PER + q];
        u32 incl = sum;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, incl, dd); if (tid >= (u32)dd) incl += o; }
        u32 run = incl - sum;
#pragma unroll
        for (u32 q = 0; q < PER; q++) { const u32 v = s.pre[tid * PER + q]; s.pre[tid * PER + q] = run; run += v; }
        if (tid == 31u) s.pre[CAP] = incl;
    }
    Team::sync();
    {
        // Many pairs (long groups): thread i takes the pairs i, i + threads, ... -- neighbouring pairs are alike in
        // length, dealt out in turn they even out over the lanes.  Few pairs: thread i takes the run
        // [P*i/threads, P*(i+1)/threads), whose pairs mostly share their first suffix.  Either way one search for
        // the place of the first pair, then on from place to place.
        const u32 P = s.pre[CAP];
        const bool in_turn = P > 5u * TT;
        const u32 pstep = in_turn ? TT : 1u;
        u32 p = in_turn ? tid : (u32)((u64)P * tid / TT);
        const u32 pend = in_turn ? P : (u32)((u64)P * (tid + 1u) / TT);
        u32 t = 0, tnext = 0, hs = 0, g = 0, pb = 0, L = 0;
        u64 xa = 0, xb = 0;
        bool busy = false;
        if (p < pend) { // the place that brings pair p = the last one whose running sum is <= p
            u32 lo = 0, hi = CAP;
            while (hi - lo > 1u) { const u32 mid = (lo + hi) >> 1; if (s.pre[mid] <= p) lo = mid; else hi = mid; }
            t = lo; tnext = s.pre[t + 1];
            hs = s.seg[t]; g = (u32)s.end[t] - hs; xa = s.x[t];
        }
        for (;;) {
            if (!busy) { // my next pair
                if (p >= pend) break;
                if (p >= tnext) {
                    do { t++; tnext = s.pre[t + 1]; } while (p >= tnext);
                    hs = s.seg[t]; g = (u32)s.end[t] - hs; xa = s.x[t];
                }
                u32 i2 = t - hs + (p - s.pre[t] + 1u);
                if (i2 >= g) i2 -= g;
                pb = hs + i2;
                xb = s.x[pb];
                L = L0;
                p += pstep;
                busy = true;
            }
            // one word of the pair (t, pb)
            if (L + 32u > Lmax) { // equal as far as the walk goes: the later one stands behind
                atomicAdd(&s.ct[t > pb ? t : pb], 0x10000u);
                busy = false;
                continue;
            }
            u64 wa, wb, d2;
            u32 ma = 0, mb = 0, dm = 0;
            if (L + 32u * WS_STEP <= Lmax) { // WS_STEP words a side at once: independent loads, one round trip to L2
                const u64 ya = xa + L, yb = xb + L;
                const u64 *pa = a.v.p2 + (ya >> 5), *pb = a.v.p2 + (yb >> 5);
                const unsigned sha = (unsigned)(ya & 31u) * 2u, shb = (unsigned)(yb & 31u) * 2u;
                u64 A[WS_STEP + 1], Bq[WS_STEP + 1];
#pragma unroll
                for (int q = 0; q <= WS_STEP; q++) { A[q] = LDG(pa + q); Bq[q] = LDG(pb + q); }
                int hit = -1;
                wa = wb = d2 = 0;
#pragma unroll
                for (int q = 0; q < WS_STEP; q++) {
                    if (hit < 0) {
                        const u64 ua = sha ? (A[q] >> sha) | (A[q + 1] << (64u - sha)) : A[q];
                        const u64 ub = shb ? (Bq[q] >> shb) | (Bq[q + 1] << (64u - shb)) : Bq[q];
                        u32 va = 0, vb = 0;
                        if (MASKS) { va = fetchm(a.v.pm, ya + 32u * q); vb = fetchm(a.v.pm, yb + 32u * q); }
                        if ((ua ^ ub) | (u64)(va ^ vb)) { hit = q; wa = ua; wb = ub; d2 = ua ^ ub; ma = va; mb = vb; dm = va ^ vb; }
                    }
                }
                if (hit < 0) { L += 32u * WS_STEP; continue; }
                L += 32u * (u32)hit;
            } else {
                wa = fetch2(a.v.p2, xa + L); wb = fetch2(a.v.p2, xb + L);
                d2 = wa ^ wb;
                if (MASKS) { ma = fetchm(a.v.pm, xa + L); mb = fetchm(a.v.pm, xb + L); dm = ma ^ mb; }
            }
            if (d2 | dm) {
                u32 f = d2 ? (u32)ctz64(d2) >> 1 : 32u;
                if (MASKS && dm) { const u32 fm = (u32)ctz32(dm); f = fm < f ? fm : f; }
                const u32 ca = (MASKS && (ma >> f & 1u)) ? 4u : (u32)(wa >> (2u * f)) & 3u;
                const u32 cb = (MASKS && (mb >> f & 1u)) ? 4u : (u32)(wb >> (2u * f)) & 3u;
                const u32 larger = ca < cb ? pb : t;
                atomicAdd(&s.ct[larger], 1u);
                atomicMax(&s.best[larger], L + f);
                busy = false;
            } else L += 32u;
        }
    }
    Team::sync();
    // new places; suffixes that are still together leave their number at the place of their first
    u32 np[WS_T], nh[WS_T];
#pragma unroll
    for (int j = 0; j < WS_T; j++) {
        if (act[j]) {
            const u32 t = tid + TT * j, c = s.ct[t] & 0xFFFFu, ti = s.ct[t] >> 16;
            nh[j] = s.seg[t] + c;
            np[j] = nh[j] + ti;
            if (ti) atomicMax(&s.clsz[nh[j]], ti + 1u);
        }
    }
    Team::sync();
#pragma unroll
    for (int j = 0; j < WS_T; j++) {
        if (act[j]) {
            const u32 t = tid + TT * j;
            a.sa[base + np[j]] = s.g[t];
            a.head[base + np[j]] = base + nh[j];
            if (ROOTS && a.roots && t == s.seg[t] && (u32)s.end[t] - t <= a.maxg) { // a walk can start here unless two are still equal
                bool ties = false;
                for (u32 q = t; q < (u32)s.end[t]; q++) ties |= s.clsz[q] != 0u;
                if (!ties) {
                    if ((u32)s.end[t] - t <= CY_MAXG) a.roots[atomicAdd(a.nroots, 1u)] = base + t;
                    else a.roots2[atomicAdd(a.nroots2, 1u)] = base + t;
                }
            }
            if (np[j] == nh[j]) {
                if (nh[j] != s.seg[t]) a.lcp[base + np[j]] = s.best[t]; // (the first of the old group keeps its border LCP)
                const u32 size = s.clsz[nh[j]];
                if (size) { // still together after Lend letters: the doubling rounds go on from there
                    a.left[atomicAdd(a.res + 0, 1u)] = ((u64)(base + nh[j]) << 32) | size;
                    atomicAdd(a.res + 1, size);
                    atomicMin(a.res + 2, Lend);
                    atomicMax(a.res + 3, size);
                }
            }
        }
    }
}

template <bool MASKS>
__global__ void __launch_bounds__(WS_WARPS * 32) k_wsort(WSortArgs a) {
    static_assert(WS_T == 4 && WS_CAP == 128, "border words are kept as two u64");
    __shared__ WsSmem<1> s_all[WS_WARPS];
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u64 r0_64 = (u64)(a.lo & ~31u) + ((u64)blockIdx.x * WS_WARPS + warp) * WS_NOM;
    if (r0_64 >= a.hi) return;
    const u32 r0 = (u32)r0_64, N = a.hi; // (heads beyond hi may be another rank's and not there yet: hi itself is a border)
    WsSmem<1> &s = s_all[warp];
    const u32 *hin = a.head; // (the carried word sort takes its groups from a list: k_wsort_list)
    // ---- the window: heads of WS_CAP places, borders as a bit set ----
    u32 hv[WS_T], bw[WS_T], sv[2];
#pragma unroll
    for (int j = 0; j < WS_T; j++) {
        const u64 p = (u64)r0 + lane + 32u * j;
        hv[j] = p < N ? (hin[p] & 0x7FFFFFFFu) : 0u;
        if (j < 2) sv[j] = p < N ? a.sa[p] : 0u; // (the suffixes of the first 64 places with the heads, not a trip later: most groups end there)
    }
#pragma unroll
    for (int j = 0; j < WS_T; j++) {
        const u64 p = (u64)r0 + lane + 32u * j;
        bw[j] = __ballot_sync(0xffffffffu, p < N ? hv[j] == (u32)p : p == N);
    }
    // starts that are this launch's: places in [lo, hi) among the first WS_NOM = 64 of the window
    static_assert(WS_NOM == 64, "a warp's own groups start in the first two border words");
    u64 mine = (u64)bw[0] | ((u64)bw[1] << 32);
    if (r0 < a.lo) mine &= ~0ull << (a.lo - r0);
    if (a.hi - r0 < 64u) mine &= (1ull << (a.hi - r0)) - 1ull;
    if (mine == 0) return; // no group starts here
    const u32 tb = (u32)__ffsll((long long)mine) - 1u;
    u32 te;
    if (a.hi - r0 < 64u) te = a.hi - r0; // hi is a border (or the end of the array)
    else if (bw[2]) te = 64u + (u32)__ffs((int)bw[2]) - 1u;
    else if (bw[3]) te = 96u + (u32)__ffs((int)bw[3]) - 1u;
    else { // the last group that starts here runs past the window: k_wsort_big's
        te = 63u - (u32)__clzll((long long)((u64)bw[0] | ((u64)bw[1] << 32)));
        // its length: gallop, then bisect (head[p] == its start for every place inside it)
        const u32 hs = r0 + te;
        u64 lo = (u64)r0 + WS_CAP - 1, step = WS_CAP; // lo is inside the group
        while (lo + step < N && (hin[lo + step] & 0x7FFFFFFFu) == hs) { lo += step; step *= 2; }
        u64 hi = lo + step < N ? lo + step : N; // first place known to be outside (or N)
        while (hi - lo > 1) {
            const u64 mid = (lo + hi) >> 1;
            if ((hin[mid] & 0x7FFFFFFFu) == hs) lo = mid; else hi = mid;
        }
        if (lane == 0) a.big[atomicAdd(a.res + 5, 1u)] = ((u64)hs << 32) | (u32)(hi - hs);
    }
    const u64 b0 = (u64)bw[0] | ((u64)bw[1] << 32), b1 = (u64)bw[2] | ((u64)bw[3] << 32);
    // ---- the suffixes of every group of two or more ----
    bool act[WS_T];
    bool any = false;
#pragma unroll
    for (int j = 0; j < WS_T; j++) {
        const u32 t = lane + 32u * j;
        act[j] = false;
        if (t >= tb && t < te) {
            const u32 hs = hv[j] - r0, he = ws_next_border(b0, b1, t);
            act[j] = he - hs >= 2u;
            s.seg[t] = (unsigned short)hs;
            s.end[t] = (unsigned short)he;
        }
        any |= act[j];
    }
    if (!__any_sync(0xffffffffu, any)) return;
    // the suffixes, where they start in the packed text, the shortest sequence of their set.  (Measured and dropped: the set
    // from the PLACE and the sequence by a search in the set's own sequence starts instead of the gathers suffix -> sequence ->
    // set -> shortest: the two searches' dependent L1 loads cost more than the gathers they save, 4.44 against 4.16 ms.)
    u32 gs[WS_T];
#pragma unroll
    for (int j = 0; j < WS_T; j++) gs[j] = !act[j] ? 0u : j < 2 ? sv[j] : a.sa[r0 + lane + 32u * j];
    u32 nmin = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < WS_T; j++) {
        if (act[j]) {
            const u32 t = lane + 32u * j;
            const u32 g = gs[j];
            const u32 k = seq_of(a.v, g);
            s.x[t] = LDG(a.v.dbl_off + k) + (g - LDG(a.v.seq_off + k));
            s.g[t] = g;
            s.ct[t] = 0; s.best[t] = 0; s.clsz[t] = 0;
            const u32 nm = LDG(a.v.seq_nmin + k);
            nmin = nm < nmin ? nm : nmin;
        }
    }
    nmin = __reduce_min_sync(0xffffffffu, nmin);
    __syncwarp();
    ws_pairs<MASKS, 1>(a, s, lane, r0, act, nmin < a.depth_cap ? nmin : a.depth_cap);
}

template <bool MASKS, bool ROOTS>
__global__ void __launch_bounds__(WS_BIG_WARPS * 32) k_wsort_big(WSortArgs a) {
    __shared__ WsSmem<WS_BIG_WARPS> s;
    const u32 tid = threadIdx.x;
    const u64 desc = a.big[blockIdx.x];
    const u32 start = (u32)(desc >> 32), size = (u32)desc;
    if (size <= a.big_min) return; // k_wsort_words'
    if (size > (u32)WS_BIG_CAP) { // longer than a CTA holds: the doubling rounds order it
        if (tid == 0) {
            a.left[atomicAdd(a.res + 0, 1u)] = desc;
            atomicAdd(a.res + 1, size);
            atomicMin(a.res + 2, a.L0);
            atomicMax(a.res + 3, size);
        }
        return;
    }
    bool act[WS_T];
#pragma unroll
    for (int j = 0; j < WS_T; j++) {
        const u32 t = tid + 32u * WS_BIG_WARPS * j;
        act[j] = t < size;
        if (act[j]) {
            const u32 g = a.sa[start + t];
            const u32 k = seq_of(a.v, g);
            s.x[t] = LDG(a.v.dbl_off + k) + (g - LDG(a.v.seq_off + k));
            s.g[t] = g;
            s.seg[t] = 0;
            s.end[t] = (unsigned short)size;
            s.ct[t] = 0; s.best[t] = 0; s.clsz[t] = 0;
        }
    }
    const u32 nmin = LDG(a.v.seq_nmin + seq_of(a.v, a.sa[start])); // a group never leaves its set
    __syncthreads();
    ws_pairs<MASKS, WS_BIG_WARPS, ROOTS>(a, s, tid, start, act, nmin < a.depth_cap ? nmin : a.depth_cap);
}

// A group of up to 256 suffixes of an ACGT-only batch, word by word instead of pair by pair: a thread per suffix, every
// round one word (32 letters) of every suffix that still shares a subgroup; a subgroup whose words differ is ranked by
// them (counting, stable) and split, the LCP at every new border read off the two words.  g loads a round where all pairs
// walk g^2/2 pairs: the roots of the carried word sort on sets of hundreds of near-identical sequences (8.4 -> 0.x ms).
// Same places, heads, LCPs, left-over list and roots as ws_pairs.
#define WSW_SMALL 256u  // a CTA of this many threads for the groups of up to as many suffixes,
#define WSW_LARGE 1024u // ... of this many for the longer ones (the same 12 letters twice in a genome)
template <bool ROOTS, u32 WSW_THREADS>
__global__ void __launch_bounds__(WSW_THREADS) k_wsort_words(WSortArgs a) {
    __shared__ u64 s_key[WSW_THREADS];
    __shared__ u32 s_flag[WSW_THREADS], s_bits[WSW_THREADS / 32u];
    const u32 tid = threadIdx.x;
    const u64 desc = a.big[blockIdx.x];
    const u32 start = (u32)(desc >> 32), g = (u32)desc;
    if (g > WSW_THREADS || (WSW_THREADS == WSW_LARGE && g <= WSW_SMALL)) return; // the other launch's (or k_wsort_big's)
    const bool have = tid < g;
    u64 x = 0;
    u32 suf = 0;
    if (have) {
        suf = a.sa[start + tid];
        const u32 k = seq_of_few(a.v, suf);
        x = LDG(a.v.dbl_off + k) + (suf - LDG(a.v.seq_off + k));
    }
    const u32 nmin = LDG(a.v.seq_nmin + seq_of_few(a.v, a.sa[start])); // a group never leaves its set
    const u32 Lmax = nmin < a.depth_cap ? nmin : a.depth_cap;
    u32 pos = tid, sub = have ? 0u : tid, end = have ? g : tid + 1u; // my place in the group, my subgroup [sub, end)
    u32 L = a.L0;
    u64 knext = (have && L + 32u <= Lmax) ? lexkey2(fetch2(a.v.p2, x + L)) : 0ull;
    bool ties = false;
    for (;; L += 32u) {
        const bool active = end - sub >= 2u;
        if (L + 32u > Lmax) { ties = __syncthreads_or(active) != 0; break; } // equal as far as the walk goes
        const u64 k = knext;
        if (active && L + 64u <= Lmax) knext = lexkey2(fetch2(a.v.p2, x + L + 32u)); // (a round ahead)
        if (active) s_key[pos] = k;
        s_flag[pos] = 0u;
        if (!__syncthreads_or(active)) break; // everybody stands alone
        const bool nonuni = active && s_key[sub] != k;
        if (!__syncthreads_or(nonuni)) continue; // every subgroup agrees on this word
        if (nonuni) s_flag[sub] = 1u;
        if (tid < WSW_THREADS / 32u) s_bits[tid] = 0u;
        __syncthreads();
        const bool split = active && s_flag[sub] != 0u;
        u32 np = pos;
        if (split) { // my place among my subgroup by this word (ties by the place before: stable)
            u32 c = 0;
            for (u32 q = sub; q < end; q++) { const u64 kq = s_key[q]; c += (kq < k || (kq == k && q < pos)) ? 1u : 0u; }
            np = sub + c;
        }
        __syncthreads();
        pos = np;
        if (split) s_key[pos] = k;
        __syncthreads();
        bool isstart = pos == sub;
        if (split && pos != sub) {
            const u64 kp = s_key[pos - 1u];
            if (kp != k) { isstart = true; a.lcp[start + pos] = L + ((u32)__clzll((long long)(kp ^ k)) >> 1); } // a border for good
        }
        if (isstart) atomicOr(&s_bits[pos >> 5], 1u << (pos & 31u));
        __syncthreads();
        { // my subgroup: the border at or before my place, the first border behind it
            u32 w = pos >> 5, m = s_bits[w] & (0xFFFFFFFFu >> (31u - (pos & 31u)));
            while (!m) m = s_bits[--w];
            sub = w * 32u + 31u - (u32)__clz((int)m);
            w = pos >> 5;
            u32 e = s_bits[w] & ((pos & 31u) < 31u ? 0xFFFFFFFFu << ((pos & 31u) + 1u) : 0u);
            while (!e && ++w < WSW_THREADS / 32u) e = s_bits[w];
            end = e ? w * 32u + (u32)__ffs((int)e) - 1u : WSW_THREADS;
        }
    }
    if (have) {
        a.sa[start + pos] = suf;
        a.head[start + pos] = start + sub;
        if (pos == sub && end - sub >= 2u) { // still together after Lend letters: the doubling rounds go on from there
            a.left[atomicAdd(a.res + 0, 1u)] = ((u64)(start + sub) << 32) | (end - sub);
            atomicAdd(a.res + 1, end - sub);
            atomicMin(a.res + 2, L);
            atomicMax(a.res + 3, end - sub);
        }
    }
    if (ROOTS && a.roots && tid == 0 && !ties && g <= a.maxg) { // a walk can start here
        if (g <= CY_MAXG) a.roots[atomicAdd(a.nroots, 1u)] = start;
        else a.roots2[atomicAdd(a.nroots2, 1u)] = start;
    }
}

// k_wsort_words for a group of up to 32 suffixes and ONE warp: a lane per suffix, the barriers are warp barriers, the
// borders one word.  s_key: 32 u64, s_flag: 32 u32, s_bits: one u32 (slices of the warp's WsSmem).
template <bool ROOTS>
__device__ __forceinline__ void ws_words_warp(const WSortArgs &a, u64 *s_key, u32 *s_flag, u32 *s_bits, const u32 lane,
                                              const u32 start, const u32 g, const u32 Lmax) {
    const bool have = lane < g;
    u64 x = 0;
    u32 suf = 0;
    if (have) {
        suf = a.sa[start + lane];
        const u32 k = seq_of_few(a.v, suf);
        x = LDG(a.v.dbl_off + k) + (suf - LDG(a.v.seq_off + k));
    }
    u32 pos = lane, sub = have ? 0u : lane, end = have ? g : lane + 1u;
    u32 L = a.L0;
    u64 knext = (have && L + 32u <= Lmax) ? lexkey2(fetch2(a.v.p2, x + L)) : 0ull;
    bool ties = false;
    for (;; L += 32u) {
        const bool active = end - sub >= 2u;
        if (L + 32u > Lmax) { ties = __any_sync(0xffffffffu, active) != 0; break; }
        const u64 k = knext;
        if (active && L + 64u <= Lmax) knext = lexkey2(fetch2(a.v.p2, x + L + 32u));
        if (active) s_key[pos] = k;
        s_flag[pos] = 0u;
        __syncwarp();
        if (!__any_sync(0xffffffffu, active)) break;
        const bool nonuni = active && s_key[sub] != k;
        const bool anysplit = __any_sync(0xffffffffu, nonuni) != 0;
        __syncwarp();
        if (!anysplit) continue;
        if (nonuni) s_flag[sub] = 1u;
        if (lane == 0) *s_bits = 0u;
        __syncwarp();
        const bool split = active && s_flag[sub] != 0u;
        u32 np = pos;
        if (split) {
            u32 c = 0;
            for (u32 q = sub; q < end; q++) { const u64 kq = s_key[q]; c += (kq < k || (kq == k && q < pos)) ? 1u : 0u; }
            np = sub + c;
        }
        __syncwarp();
        pos = np;
        if (split) s_key[pos] = k;
        __syncwarp();
        bool isstart = pos == sub;
        if (split && pos != sub) {
            const u64 kp = s_key[pos - 1u];
            if (kp != k) { isstart = true; a.lcp[start + pos] = L + ((u32)__clzll((long long)(kp ^ k)) >> 1); }
        }
        if (isstart) atomicOr(s_bits, 1u << pos);
        __syncwarp();
        {
            const u32 b = *s_bits;
            sub = 31u - (u32)__clz((int)(b & (0xFFFFFFFFu >> (31u - pos))));
            const u32 e = b & (pos < 31u ? 0xFFFFFFFFu << (pos + 1u) : 0u);
            end = e ? (u32)__ffs((int)e) - 1u : 32u;
        }
        __syncwarp();
    }
    if (have) {
        a.sa[start + pos] = suf;
        a.head[start + pos] = start + sub;
        if (pos == sub && end - sub >= 2u) {
            a.left[atomicAdd(a.res + 0, 1u)] = ((u64)(start + sub) << 32) | (end - sub);
            atomicAdd(a.res + 1, end - sub);
            atomicMin(a.res + 2, L);
            atomicMax(a.res + 3, end - sub);
        }
    }
    if (ROOTS && a.roots && lane == 0 && !ties && g <= a.maxg) a.roots[atomicAdd(a.nroots, 1u)] = start; // (g <= 32 = CY_MAXG)
}

// the groups of a list (k_cylist), one warp a group, the warps fetching entries until the list is done
template <bool MASKS>
__global__ void __launch_bounds__(WSL_WARPS * 32) k_wsort_list(WSortArgs a, const u64 *list, const u32 *count, u32 *next) {
    __shared__ WsSmem<1> s_all[WSL_WARPS];
    const u32 lane = threadIdx.x & 31u;
    WsSmem<1> &s = s_all[threadIdx.x >> 5];
    const u32 n = *count;
    for (;;) {
        u32 i = 0;
        if (lane == 0) i = atomicAdd(next, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) return;
        const u64 desc = list[i];
        const u32 start = (u32)(desc >> 32), size = (u32)desc;
        if (size > (a.list_cap ? a.list_cap : (u32)WS_CAP)) { if (lane == 0) a.big[atomicAdd(a.res + 5, 1u)] = desc; continue; }
        if (!MASKS && a.words_small && size <= 32u) {
            const u32 nm = LDG(a.v.seq_nmin + seq_of_few(a.v, a.sa[start])); // a group never leaves its set
            ws_words_warp<true>(a, s.x, s.ct, s.best, lane, start, size, nm < a.depth_cap ? nm : a.depth_cap);
            __syncwarp();
            continue;
        }
        bool act[WS_T];
        u32 nmin = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < WS_T; j++) {
            const u32 t = lane + 32u * j;
            act[j] = t < size;
            if (act[j]) {
                const u32 g = a.sa[start + t];
                const u32 k = seq_of_few(a.v, g);
                s.x[t] = LDG(a.v.dbl_off + k) + (g - LDG(a.v.seq_off + k));
                s.g[t] = g;
                s.seg[t] = 0;
                s.end[t] = (unsigned short)size;
                s.ct[t] = 0; s.best[t] = 0; s.clsz[t] = 0;
                if (t == 0) nmin = LDG(a.v.seq_nmin + k); // a group never leaves its set
            }
        }
        nmin = __shfl_sync(0xffffffffu, nmin, 0);
        __syncwarp();
        ws_pairs<MASKS, 1, true>(a, s, lane, start, act, nmin < a.depth_cap ? nmin : a.depth_cap);
        __syncwarp();
    }
}
static inline void launch_wsort_list(Exec &ex, const WSortArgs &a, const u64 *list, const u32 *count) {
    PROF_BEGIN(ex, a.roots ? "k_wsort_list(roots)" : "k_wsort_list(rest)", 0.0);
    if (a.masks) k_wsort_list<true><<<148 * 4, WSL_WARPS * 32, 0, ex.stream>>>(a, list, count, const_cast<u32 *>(count) + 1);
    else k_wsort_list<false><<<148 * 4, WSL_WARPS * 32, 0, ex.stream>>>(a, list, count, const_cast<u32 *>(count) + 1);
    PROF_END(ex);
    ex.launches++;
}
static inline void launch_wsort(Exec &ex, const WSortArgs &a) {
    if (a.hi <= a.lo) return;
    const u32 nwarps = (a.hi - (a.lo & ~31u) + WS_NOM - 1) / WS_NOM;
    PROF_BEGIN(ex, "k_wsort", 4.0 * a.N);
    if (a.masks) k_wsort<true><<<(nwarps + WS_WARPS - 1) / WS_WARPS, WS_WARPS * 32, 0, ex.stream>>>(a);
    else k_wsort<false><<<(nwarps + WS_WARPS - 1) / WS_WARPS, WS_WARPS * 32, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
static inline void launch_wsort_big(Exec &ex, const WSortArgs &a0) {
    if (a0.nbig == 0) return;
    WSortArgs a = a0;
    if (!a.masks) { // ACGT only: the groups of up to 1024 word by word
        a.big_min = WSW_LARGE;
        PROF_BEGIN(ex, a.roots ? "k_wsort_words(roots)" : "k_wsort_words", 0.0);
        if (a.roots) { k_wsort_words<true, WSW_SMALL><<<a.nbig, WSW_SMALL, 0, ex.stream>>>(a); k_wsort_words<true, WSW_LARGE><<<a.nbig, WSW_LARGE, 0, ex.stream>>>(a); }
        else { k_wsort_words<false, WSW_SMALL><<<a.nbig, WSW_SMALL, 0, ex.stream>>>(a); k_wsort_words<false, WSW_LARGE><<<a.nbig, WSW_LARGE, 0, ex.stream>>>(a); }
        PROF_END(ex);
        ex.launches += 2;
    }
    PROF_BEGIN(ex, a.roots ? "k_wsort_big(roots)" : "k_wsort_big", 0.0);
    if (a.roots) { // (the carried word sort's roots of more than a warp's window: the walks start from them too)
        if (a.masks) k_wsort_big<true, true><<<a.nbig, WS_BIG_WARPS * 32, 0, ex.stream>>>(a);
        else k_wsort_big<false, true><<<a.nbig, WS_BIG_WARPS * 32, 0, ex.stream>>>(a);
    } else if (a.masks) k_wsort_big<true, false><<<a.nbig, WS_BIG_WARPS * 32, 0, ex.stream>>>(a);
    else k_wsort_big<false, false><<<a.nbig, WS_BIG_WARPS * 32, 0, ex.stream>>>(a);
    PROF_END(ex);
    ex.launches++;
}
#endif

// LCP of the places the word sort could not give (groups finished by the doubling rounds)
struct LcpFixArgs { BatchView v; const u32 *sa; u32 *lcp; const u32 *any_other; };
HD void lcpfix_body(long long i, const LcpFixArgs &a) {
    if (a.lcp[i] != LCP_UNKNOWN) return;
    LcpDirectArgs d{a.v, a.sa, a.lcp, a.any_other, 1, nullptr};
    lcpdirect_body(i, d);
}
MAP_KERNEL(lcpfix, LcpFixArgs, 4)
