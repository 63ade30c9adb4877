// rare.cuh -- the sets on which the reference's TREE differs from a plain suffix array: a suffix that
// matches its neighbour over a whole sequence length (lcp >= the set's shortest sequence).  Two causes:
//
//   * a sequence that is a power w^c: its c identical rotations share ONE leaf of the tree, whose
//     `rotation` stays the first one (gencycsuffixtrees.c:507-517);
//   * a whole rotation of the shortest sequence occurs in every other sequence: a LEAF that holds all
//     sequences.  Leaves carry the link to the NEXT ROTATION's leaf where inner nodes carry the suffix
//     link (gencycsuffixtrees.c:505), and removeSuffixNodes (csamsa.c:80) / collectNodeChains
//     (csamsa.c:147-183) follow it: the list surgery of csamsa.c:88-101 then depends on the order of the
//     list, and the walk of csamsa.c:153 leaves the tree.
//
// k_leafscan marks such sets (one streaming pass over the LCP array, nothing else on the normal path).
// For a marked set -- there are none in a batch of ordinary genomes -- the three kernels below redo
// the steps concerned the way the reference does them, ONE THREAD PER SET (the reference's list walk
// is a chain of dependent steps; speed is not the point here, equality with the reference is).  They are
// launched for every batch and return at once for a set that is not marked: the host never waits to
// find out whether there is one.
#pragma once

#define CSA_FLAG_DEGENERATE 1u  // collectNodeChains walks off a leaf (csamsa.c:153)
#define CSA_FLAG_HANG 2u        // csamsa.c:197 never returns
#define CSA_FLAG_RARE 4u        // a full-length match somewhere in the set: the kernels of this file run
#define CSA_FLAG_UNDEFINED 8u   // removeSuffixNodes frees the list item it stands on

struct LeafScanArgs { BatchView v; const u32 *lcp; u32 *set_flags; u32 batch_nmin; u32 off; };
HD void leafscan_body(long long i0, const LeafScanArgs &a) {
    const u32 i = (u32)i0 + a.off;
    u32 l = a.lcp[i];
    if (l < a.batch_nmin || l == 0xFFFFFFFFu) return; // (no set's shortest sequence is shorter: no search for the set)
    u32 s = set_of_pos(a.v, (u32)i);
    if ((u32)i == LDG(a.v.set_base0 + s)) return;     // the first place of a set has no left neighbour
    if (l >= LDG(a.v.set_nmin + s)) ATOMIC_OR(a.set_flags + s, CSA_FLAG_RARE);
}
#ifdef CSA_EMU
MAP_KERNEL(leafscan, LeafScanArgs, 4)
#else
// four places a thread, their LCPs in one 16-byte load (a 4-byte stream a thread ran at 2 TB/s); almost no place passes
__global__ void __launch_bounds__(256) k_leafscan(long long n, LeafScanArgs a) {
    const long long i0 = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (i0 >= n) return;
    if (i0 + 4 <= n && ((a.off + (u32)i0) & 3u) == 0u) {
        const uint4 v = *reinterpret_cast<const uint4 *>(a.lcp + a.off + i0);
        const u32 l[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) if (l[q] >= a.batch_nmin && l[q] != 0xFFFFFFFFu) leafscan_body(i0 + q, a);
    } else for (long long i = i0; i < n && i < i0 + 4; i++) leafscan_body(i, a);
}
static inline void launch_leafscan(Exec &ex, long long n, LeafScanArgs a) {
    if (n <= 0) return;
    PROF_BEGIN(ex, "k_leafscan", 4.0 * n);
    k_leafscan<<<(unsigned)((n + 1023) / 1024), 256, 0, ex.stream>>>(n, a);
    PROF_END(ex);
    ex.launches++;
}
#endif

HD u32 rare_seq_len(const BatchView &v, u32 k) { return LDG(v.seq_off + k + 1) - LDG(v.seq_off + k); }

// ---- 1. identical rotations of one sequence: one leaf -------------------------------------------------
// The copies (k,p+|w|), (k,p+2|w|), ... stand right behind (k,p) in the suffix array (equal strings, ties by
// index).  They move to the end of the set's range with lcp 0 -- there they are leaves under the root that
// no window with a depth can hold -- and set_neff[s] counts what is left in front.
struct RareCollapseArgs { BatchView v; u32 *sa; u32 *lcp; const u32 *set_flags; u32 *scratch; u32 *set_neff; u32 *seq_per; };
HD void rarecollapse_body(long long si, const RareCollapseArgs &a) {
    const u32 s = (u32)si;
    const u32 s0 = LDG(a.v.set_base0 + s), s1 = LDG(a.v.set_base0 + s + 1);
    a.set_neff[s] = s1 - s0;
    if (!(a.set_flags[s] & CSA_FLAG_RARE)) return;
    for (u32 k = LDG(a.v.set_seq0 + s); k < LDG(a.v.set_seq0 + s + 1); k++) a.seq_per[k] = rare_seq_len(a.v, k);
    u32 w = s0, nd = 0, run = 0xFFFFFFFFu, prevk = CSA_NONE;
    for (u32 i = s0; i < s1; i++) {
        const u32 g = a.sa[i], k = seq_of(a.v, g), l = (i == s0) ? 0u : a.lcp[i];
        if (l < run) run = l;
        const bool dup = i > s0 && prevk == k && l >= rare_seq_len(a.v, k);
        prevk = k;
        if (dup) {
            const u32 p = g - LDG(a.v.seq_off + k);
            if (p < a.seq_per[k]) a.seq_per[k] = p;
            a.scratch[s0 + nd++] = g;
            continue;
        }
        a.sa[w] = g;
        if (w != s0) a.lcp[w] = run;
        w++;
        run = 0xFFFFFFFFu;
    }
    for (u32 j = 0; j < nd; j++) { a.sa[w + j] = a.scratch[s0 + j]; a.lcp[w + j] = 0; }
    a.set_neff[s] = w - s0;
}
MAP_KERNEL(rarecollapse, RareCollapseArgs, 0)

// ---- 2. collectNodes + removeSuffixNodes + removeNonUniqueNodes, the reference's way ----------------------
struct RareBlocksArgs {
    BatchView v; const u32 *sa; const u32 *lcp; u32 *R; int have_R; // have_R == 0: the cover array was not built (k_blockfind2's path): made here for the set
    u32 *set_flags; const u32 *set_neff; const u32 *seq_per;
    u32 *isa;                                    // [N] written here for the set's range
    Seq0Q q;                                     // sequence 0 of every set: which of two leaves the DFS meets first
    u32 *w[6];                                   // scratch, N entries each; a set uses its own range
    u32 *isblock; u32 *depth;
    u32 *rare_collected; u32 *rare_suffixfree;
};
// smallest / largest place of the LCP interval of the first `len` letters of the suffix at place r
HD void rare_span(const u32 *lcp, u32 lo, u32 hi, u32 r, u32 len, u32 *l_out, u32 *r_out) {
    u32 l = r;
    while (l > lo && lcp[l] >= len) l--;
    while (r + 1 < hi && lcp[r + 1] >= len) r++;
    *l_out = l; *r_out = r;
}
// heap sort of idx[0..n), DESCENDING by `less` (a min-heap emptied towards the end)
template <class Less> HD void rare_sift(const Less &less, u32 *idx, u32 root, u32 n) {
    for (;;) {
        u32 c = 2 * root + 1;
        if (c >= n) return;
        if (c + 1 < n && less(idx[c + 1], idx[c])) c++;
        if (!less(idx[c], idx[root])) return;
        u32 t = idx[root]; idx[root] = idx[c]; idx[c] = t;
        root = c;
    }
}
template <class Less> HD void rare_sort_desc(const Less &less, u32 *idx, u32 n) {
    for (u32 i = 0; i < n; i++) idx[i] = i;
    for (u32 i = n / 2; i-- > 0;) rare_sift(less, idx, i, n);
    for (u32 x = n; x > 1; x--) { u32 t = idx[0]; idx[0] = idx[x - 1]; idx[x - 1] = t; rare_sift(less, idx, 0, x - 1); }
}
// list order of two nodes (nodeslinkedlists.c:36): deeper first, of equal depth the one the DFS meets LATER first
struct RareNodeLess {
    const u32 *depth; const u32 *leaf; const Seq0Q *q;
    HDM bool operator()(u32 x, u32 y) const { // x stands BEHIND y in the list ("smaller")
        if (depth[x] != depth[y]) return depth[x] < depth[y];
        return seq0_before(*q, leaf[x], leaf[y]);
    }
};
struct RarePairLess {
    const u32 *hi; const u32 *lo;
    HDM bool operator()(u32 x, u32 y) const { return hi[x] != hi[y] ? hi[x] < hi[y] : lo[x] < lo[y]; }
};
HD void rareblocks_body(long long si, const RareBlocksArgs &a) {
    const u32 s = (u32)si;
    if (!(a.set_flags[s] & CSA_FLAG_RARE)) return;
    const u32 s0 = LDG(a.v.set_base0 + s), s1 = LDG(a.v.set_base0 + s + 1), end = s0 + a.set_neff[s];
    const u32 m = LDG(a.v.set_seq0 + s + 1) - LDG(a.v.set_seq0 + s);
    const u32 nmin = LDG(a.v.set_nmin + s);
    const u32 L = s1 - s0, H = (L + 1) / 2; // a set has at most L/2 collected nodes (each holds >= m >= 2 places)
    for (u32 i = s0; i < s1; i++) { a.isa[a.sa[i]] = i; a.isblock[i] = 0; }
    if (!a.have_R) { // R[l] = smallest r such that [l, r] holds every sequence of the set (two pointers; counts in the depth slice, free until the end)
        const u32 q0 = LDG(a.v.set_seq0 + s);
        u32 *cnt = a.depth + s0;
        for (u32 k = 0; k < m; k++) cnt[k] = 0;
        u32 covered = 0, r = s0;
        for (u32 l = s0; l < s1; l++) {
            while (covered < m && r < end) { if (cnt[seq_of(a.v, a.sa[r]) - q0]++ == 0) covered++; r++; }
            a.R[l] = (covered == m && l < end) ? r - 1 : s1;
            if (l < end && --cnt[seq_of(a.v, a.sa[l]) - q0] == 0) covered--;
        }
    }
    // -- collectNodes (csamsa.c:64): the all-sequence LCP intervals without an all-sequence child
    u32 *st_lcp = a.w[0] + s0, *st_lb = a.w[1] + s0, *st_child = a.w[2] + s0; // stack: at most L deep
    u32 *c_lb = a.w[3] + s0, *c_rb = a.w[4] + s0, *c_depth = a.w[5] + s0;     // lower halves of their slices
    u32 C = 0, top = 0;
    st_lcp[0] = 0; st_lb[0] = s0; st_child[0] = 0; top = 1;
    for (u32 i = s0 + 1; i <= end && top > 0; i++) {
        const long long cur = (i < end) ? (long long)a.lcp[i] : -1;
        u32 lb = i - 1, carry_all = 0;
        while (top > 0 && cur < (long long)st_lcp[top - 1]) {
            if (carry_all) st_child[top - 1] = 1;
            const u32 nlb = st_lb[top - 1];
            const u32 all = a.R[nlb] <= i - 1 ? 1u : 0u;
            if (all && !st_child[top - 1]) { c_lb[C] = nlb; c_rb[C] = i - 1; c_depth[C] = st_lcp[top - 1]; C++; }
            carry_all = all;
            lb = nlb;
            top--;
        }
        if (top == 0) break;
        if (cur > (long long)st_lcp[top - 1]) { st_lcp[top] = (u32)cur; st_lb[top] = lb; st_child[top] = carry_all; top++; }
        else if (carry_all) st_child[top - 1] = 1;
    }
    a.rare_collected[s] = C;
    a.rare_suffixfree[s] = C;
    if (C == 0) return;
    // -- list order (nodeslinkedlists.c:36): depth descending, then the node met LATER in the DFS first
    // (the stack is done, its arrays are free again)
    u32 *c_leaf = st_lcp, *ord = st_child, *nxt = c_lb + H, *prv = c_rb + H, *gone = c_depth + H;
    (void)st_lb;
    {
        // (the first leaf of sequence 0 at or behind a node's left border lies below the node)
        for (u32 c = 0; c < C; c++) c_leaf[c] = seq0_leaf_at(a.q, s, c_lb[c]);
        const RareNodeLess less{c_depth, c_leaf, &a.q};
        rare_sort_desc(less, ord, C);
    }
    // ord[i] = node at list place i
    for (u32 i = 0; i < C; i++) { nxt[i] = i + 1 < C ? i + 1 : CSA_NONE; prv[i] = i ? i - 1 : CSA_NONE; gone[i] = 0; }
    const bool root_only = c_depth[ord[0]] == 0; // csamsa.c:85: a list that holds only the root is left alone
    // -- removeSuffixNodes (csamsa.c:80-110), the list walk itself
    u32 left = C;
    bool undefined = false;
    for (u32 node = root_only ? CSA_NONE : 0u; node != CSA_NONE && !undefined; node = nxt[node]) {
        u32 g = a.sa[c_lb[ord[node]]], len = c_depth[ord[node]];
        u32 search = nxt[node];
        for (;;) {
            {   // getSuffixNode (gencycsuffixtrees.c:327): the string loses its first letter
                const u32 k = seq_of(a.v, g), off = LDG(a.v.seq_off + k);
                u32 p = g - off + 1;
                if (p >= a.seq_per[k]) p -= a.seq_per[k];
                g = off + p;
            }
            if (len <= 1 || search == CSA_NONE) break; // csamsa.c:92 `searchnode!=NULL && suffix!=root`
            len--;
            u32 l, r, d;
            rare_span(a.lcp, s0, end, a.isa[g], len, &l, &r);
            if (l == r) d = rare_seq_len(a.v, seq_of(a.v, a.sa[l]));
            else { d = 0xFFFFFFFFu; for (u32 q = l + 1; q <= r; q++) if (a.lcp[q] < d) d = a.lcp[q]; }
            while (search != CSA_NONE && d < c_depth[ord[search]]) search = nxt[search];
            while (search != CSA_NONE && d == c_depth[ord[search]]) {
                if (c_lb[ord[search]] == l && c_rb[ord[search]] == r) {
                    const u32 del = search, p = prv[del], nx = nxt[del];
                    if (nx != CSA_NONE) prv[nx] = p;
                    if (p != CSA_NONE) nxt[p] = nx;
                    search = p != CSA_NONE ? p : nx; // nodeslinkedlists.c:88 deleteItem: the item before, else the one behind
                    gone[del] = 1;
                    left--;
                    if (del == node) undefined = true; // the reference goes on reading the freed item
                    break;
                }
                search = nxt[search];
            }
        }
    }
    a.rare_suffixfree[s] = left;
    if (undefined) { ATOMIC_OR(a.set_flags + s, CSA_FLAG_UNDEFINED); return; }
    // -- removeNonUniqueNodes (csamsa.c:283): one rotation of every sequence
    for (u32 i = 0; i < C; i++) {
        const u32 c = ord[i];
        if (gone[i] || c_rb[c] - c_lb[c] + 1 != m) continue;
        a.isblock[c_lb[c]] = 1;
        a.depth[c_lb[c]] = c_depth[c];
    }
    (void)nmin;
}
MAP_KERNEL(rareblocks, RareBlocksArgs, 0)

// ---- 3. the walk of csamsa.c:147-183 for a marked set -----------------------------------------------
// Replaces what k_link found for the set's blocks: succ_lo == succ_hi == the one block that follows in every
// sequence, or a pair that differs (k_gap then reads "no successor").
struct RareWalkArgs {
    BatchView v; const u32 *sa; const u32 *lcp; const u32 *R; u32 *set_flags; const u32 *set_neff; const u32 *seq_per;
    const u32 *isa; const u32 *set_blk0; const u32 *set_pos0; const u32 *o_depth; const int *o_pos;
    u32 *w[6];
    u32 *succ_lo; u32 *succ_hi;
};
HD void rarewalk_body(long long si, const RareWalkArgs &a) {
    const u32 s = (u32)si;
    if (!(a.set_flags[s] & CSA_FLAG_RARE) || (a.set_flags[s] & CSA_FLAG_UNDEFINED)) return;
    const u32 s0 = LDG(a.v.set_base0 + s), end = s0 + a.set_neff[s];
    const u32 q0 = LDG(a.v.set_seq0 + s), m = LDG(a.v.set_seq0 + s + 1) - q0;
    const u32 nmin = LDG(a.v.set_nmin + s);
    const u32 b0 = a.set_blk0[s], B = a.set_blk0[s + 1] - b0, po = a.set_pos0[s];
    if (B == 0) return;
    // (B blocks <= the leaves of any sequence; a sequence meets each block n/per times a lap: <= n events)
    u32 *ev_e = a.w[0] + s0, *ev_b = a.w[1] + s0, *ord = a.w[2] + s0, *nextv = a.w[3] + s0, *invalid = a.w[4] + s0;
    for (u32 b = 0; b < B; b++) { nextv[b] = CSA_NONE; invalid[b] = 0; }
    bool bad = false;
    for (u32 k = 0; k < m && !bad; k++) {
        const u32 off = LDG(a.v.seq_off + q0 + k), n = rare_seq_len(a.v, q0 + k), per = a.seq_per[q0 + k], reps = n / per;
        // first place of the text where the walk stands on a leaf that holds every sequence: the nmin letters
        // before it are a whole rotation of the shortest sequence and occur in all
        u32 bad_e = 0xFFFFFFFFu;
        for (u32 p = 0; p < per && bad_e == 0xFFFFFFFFu; p++) {
            u32 l, r;
            rare_span(a.lcp, s0, end, a.isa[off + p], nmin, &l, &r);
            if (r > l && a.R[l] <= r) bad_e = p + nmin;
        }
        u32 ne = 0;
        for (u32 b = 0; b < B; b++)
            for (u32 t = 0; t < reps; t++) {
                const u32 e = (u32)a.o_pos[po + b * m + k] + t * per + a.o_depth[b0 + b];
                ev_e[ne] = e; ev_b[ne] = b;
                ne++;
            }
        { const RarePairLess less{ev_e, ev_b}; rare_sort_desc(less, ord, ne); } // read backwards: ascending by (place, block)
        u32 limit = n, prev = CSA_NONE;
        bool first = true;
        for (u32 x = ne; x-- > 0;) {
            const u32 e = ev_e[ord[x]], b = ev_b[ord[x]];
            if (bad_e <= e && bad_e < limit) { bad = true; break; } // (the leaf is met before a block noticed at the same place)
            if (e >= limit) break;
            if (first) { limit = n + (e - a.o_depth[b0 + b]); first = false; } // csamsa.c:168
            if (prev != CSA_NONE && !invalid[prev]) {
                if (nextv[prev] == CSA_NONE) nextv[prev] = b;
                else if (nextv[prev] != b) invalid[prev] = 1;
            }
            prev = b;
        }
        if (!bad && bad_e < limit) bad = true;
    }
    if (bad) { ATOMIC_OR(a.set_flags + s, CSA_FLAG_DEGENERATE); return; }
    for (u32 b = 0; b < B; b++) {
        if (invalid[b] || nextv[b] == CSA_NONE) { a.succ_lo[b0 + b] = CSA_NONE; a.succ_hi[b0 + b] = 0; }
        else { a.succ_lo[b0 + b] = b0 + nextv[b]; a.succ_hi[b0 + b] = b0 + nextv[b]; }
    }
}
MAP_KERNEL(rarewalk, RareWalkArgs, 0)
