/*
 * csa_oracle.c -- CPU restatement of fjdf/CSA's `./CSA R` rotation path.
 * TEST INFRASTRUCTURE ONLY (see csa_oracle.h).  Parity status: PINNED against the
 * compiled reference (oracle/validate_against_ref.py) and tests/golden/.
 *
 * How the reference's suffix-tree vocabulary maps onto a suffix array:
 *
 *   tree                                            | here
 *   ------------------------------------------------+---------------------------------------
 *   compact trie of every rotation of every         | all (k,p) sorted by the cyclic string
 *   sequence  gencycsuffixtrees.c:418-545            | s_k[p..] (prefix doubling) + LCP
 *   internal node                                    | LCP interval [lb,rb], depth = its lcp
 *   node->fromseqs == allseqsmask  (:34)             | every sequence occurs in SA[lb..rb]
 *   collectNodes (csamsa.c:64): deepest all-seq      | all-seq interval with no all-seq child
 *   removeSuffixNodes (csamsa.c:80)                  | drop X when some x.X is common to all
 *   removeNonUniqueNodes (csamsa.c:283)              | keep intervals of exactly m suffixes
 *   order of blockslist: insertSortedItem            | depth desc, then LATER-visited first;
 *   (nodeslinkedlists.c:36) over a DFS whose child   | DFS order == order of first occurrence
 *   order is creation order (addBranch :193)         | in sequence 0 (see dfs_rank_seq0)
 *   collectNodeChains (csamsa.c:135)                 | chain_blocks() below, literal
 *   sortList (nodeslinkedlists.c:59)                 | stable sort by size, descending
 */
#include "csa_oracle.h"
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long long u64;

/* gencycsuffixtrees.c:283/297: every letter that is not A/C/G/T is the same 5th letter */
static inline unsigned char code_of(char c) {
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 4;
    }
}

typedef struct gtext {
    int m;
    int N;               /* total number of rotations = sum of lengths */
    int nmax, nmin;
    int *off;            /* m+1 */
    const int *n;        /* lengths */
    unsigned char *code; /* N */
    int *seqof;          /* N */
    int *per;            /* m: smallest d with s_k rotated by d == s_k (n[k] when s_k is not a power w^c) */
} gtext;

static inline int cyc(const gtext *t, int g, int h) {
    int k = t->seqof[g];
    int p = g - t->off[k];
    long long q = ((long long)p + h) % t->n[k];
    return t->off[k] + (int)q;
}

static int gtext_init(gtext *t, int m, const char *const *texts, const int *sizes) {
    long long tot = 0;
    t->m = m;
    t->n = sizes;
    t->off = (int *)malloc(sizeof(int) * (m + 1));
    t->nmax = 0;
    t->nmin = INT_MAX;
    for (int k = 0; k < m; k++) {
        t->off[k] = (int)tot;
        tot += sizes[k];
        if (sizes[k] > t->nmax) t->nmax = sizes[k];
        if (sizes[k] < t->nmin) t->nmin = sizes[k];
    }
    if (tot >= INT_MAX) return -1;
    t->off[m] = (int)tot;
    t->N = (int)tot;
    t->code = (unsigned char *)malloc(t->N ? t->N : 1);
    t->seqof = (int *)malloc(sizeof(int) * (t->N ? t->N : 1));
    for (int k = 0; k < m; k++)
        for (int p = 0; p < sizes[k]; p++) {
            t->code[t->off[k] + p] = code_of(texts[k][p]);
            t->seqof[t->off[k] + p] = k;
        }
    /* gencycsuffixtrees.c:507-517: when the leaf of a rotation already exists and belongs to the
     * sequence being inserted (s_k = w^c: rotation p+|w| spells the same n letters as rotation p) nothing
     * is added -- the c identical rotations share ONE leaf, whose `rotation` stays the first p. */
    t->per = (int *)malloc(sizeof(int) * m);
    for (int k = 0; k < m; k++) {
        int n = sizes[k], *f = (int *)malloc(sizeof(int) * (n + 1));
        const unsigned char *c = t->code + t->off[k];
        f[0] = -1; /* KMP failure function: n - f[n] is the period when it divides n */
        for (int i = 0, j = -1; i < n;) {
            while (j >= 0 && c[i] != c[j]) j = f[j];
            i++; j++;
            f[i] = j;
        }
        int d = n - f[n];
        t->per[k] = (n % d == 0) ? d : n;
        free(f);
    }
    return 0;
}

static void gtext_free(gtext *t) {
    free(t->off);
    free(t->code);
    free(t->seqof);
    free(t->per);
}

/* stable counting sort of idx[0..N) by key[idx[i]] in [0,K) */
static void counting_sort(const int *in, int *out, int N, const int *key, int K, int *cnt) {
    memset(cnt, 0, sizeof(int) * (size_t)(K + 1));
    for (int i = 0; i < N; i++) cnt[key[in[i]] + 1]++;
    for (int i = 0; i < K; i++) cnt[i + 1] += cnt[i];
    for (int i = 0; i < N; i++) out[cnt[key[in[i]]]++] = in[i];
}

/* Generalized cyclic suffix array: every rotation (k,p) sorted by the periodic string
 * s_k[p], s_k[p+1], ...  Replaces Ukkonen's construction at gencycsuffixtrees.c:418. */
static void build_gsa(const gtext *t, int *sa, int *isa) {
    int N = t->N;
    int *rank = (int *)calloc(N + 1, sizeof(int));
    int *key2 = (int *)malloc(sizeof(int) * N);
    int *tmp = (int *)calloc(N + 1, sizeof(int));
    int *cnt = (int *)malloc(sizeof(int) * ((size_t)N + 8));
    for (int g = 0; g < N; g++) {
        rank[g] = t->code[g];
        tmp[g] = g;
    }
    counting_sort(tmp, sa, N, rank, 5, cnt);
    /* densify initial ranks */
    {
        int r = -1, prev = -1;
        for (int i = 0; i < N; i++) {
            int c = t->code[sa[i]];
            if (c != prev) { r++; prev = c; }
            key2[sa[i]] = r;
        }
        memcpy(rank, key2, sizeof(int) * N);
    }
    int nranks = 0;
    for (int g = 0; g < N; g++) if (rank[g] + 1 > nranks) nranks = rank[g] + 1;
    for (long long h = 1; nranks < N && h < 2LL * t->nmax; h *= 2) {
        for (int g = 0; g < N; g++) key2[g] = rank[cyc(t, g, (int)(h % INT_MAX))];
        counting_sort(sa, tmp, N, key2, nranks, cnt);
        counting_sort(tmp, sa, N, rank, nranks, cnt);
        int r = 0;
        tmp[sa[0]] = 0;
        for (int i = 1; i < N; i++) {
            if (rank[sa[i]] != rank[sa[i - 1]] || key2[sa[i]] != key2[sa[i - 1]]) r++;
            tmp[sa[i]] = r;
        }
        memcpy(rank, tmp, sizeof(int) * N);
        nranks = r + 1;
    }
    for (int i = 0; i < N; i++) isa[sa[i]] = i;
    free(rank); free(key2); free(tmp); free(cnt);
}

/* lcp[i] = length of the common prefix of rotations sa[i-1], sa[i], never more than either
 * rotation's length (a tree path ends at depth textsize, gencycsuffixtrees.c:500). */
static void build_lcp(const gtext *t, const int *sa, const int *isa, int *lcp) {
    for (int k = 0; k < t->m; k++) {
        int h = 0;
        for (int p = 0; p < t->n[k]; p++) {
            int g = t->off[k] + p;
            int r = isa[g];
            if (r == 0) { lcp[0] = 0; h = 0; continue; }
            int b = sa[r - 1];
            int cap = t->n[k] < t->n[t->seqof[b]] ? t->n[k] : t->n[t->seqof[b]];
            if (h > cap) h = cap;
            while (h < cap && t->code[cyc(t, g, h)] == t->code[cyc(t, b, h)]) h++;
            lcp[r] = h;
            if (h > 0) h--;
        }
    }
}

/* ---- interval scan: collectNodes / removeSuffixNodes / removeNonUniqueNodes ---------- */

typedef struct blk {
    int depth, lb;
    int dfs;     /* DFS visit index of the block (first-occurrence order in sequence 0) */
    int *pos;    /* m positions */
} blk;

typedef struct cnode { /* a node collectNodes (csamsa.c:64) puts on blockslist */
    int lb, rb, depth;
    int dfs;   /* DFS visit index (dfs_rank_seq0 of a sequence-0 rotation below the node) */
    int gone;  /* removeSuffixNodes deleted it */
    int sfx;   /* some x.X is common to all sequences */
} cnode;

typedef struct scan_out {
    int count_collected, count_suffixfree, count_unique;
    int undefined; /* removeSuffixNodes frees the list item it stands on (see remove_suffix_nodes) */
    cnode *nodes;
    int nnodes, ncap;
    blk *blocks;
    int nblocks, cap;
} scan_out;

#define NMASK 6 /* [0] sequences present, [1..5] sequences with an occurrence preceded by letter x */

/* csamsa.c:64 collectNodes: the all-sequence LCP intervals without an all-sequence child */
static void scan_intervals(const gtext *t, const int *sa, const int *lcp, int N, scan_out *o) {
    int m = t->m, W = (m + 63) / 64;
    size_t ew = (size_t)NMASK * W;
    int cap = 1024, top = 0;
    int *s_lcp = (int *)malloc(sizeof(int) * cap), *s_lb = (int *)malloc(sizeof(int) * cap);
    char *s_allchild = (char *)malloc(cap);
    u64 *s_mask = (u64 *)malloc(sizeof(u64) * ew * cap);
    u64 *carry = (u64 *)malloc(sizeof(u64) * ew);
    u64 *full = (u64 *)calloc(W, sizeof(u64));
    for (int k = 0; k < m; k++) full[k >> 6] |= 1ULL << (k & 63);
    memset(o, 0, sizeof(*o));
    /* root */
    s_lcp[0] = 0; s_lb[0] = 0; s_allchild[0] = 0;
    memset(s_mask, 0, sizeof(u64) * ew);
    top = 1;
    for (int i = 1; i <= N; i++) {
        int cur = (i < N) ? lcp[i] : -1;
        int lb = i - 1;
        /* leaf i-1 (one rotation): gencycsuffixtrees.c:160 newNode fromseqs=masks[currentseq] */
        int g = sa[i - 1], k = t->seqof[g];
        int x = t->code[cyc(t, g, t->n[k] - 1)];
        int carry_allseq = 0;
        memset(carry, 0, sizeof(u64) * ew);
        carry[k >> 6] |= 1ULL << (k & 63);
        carry[(size_t)(1 + x) * W + (k >> 6)] |= 1ULL << (k & 63);
        while (top > 0 && cur < s_lcp[top - 1]) {
            u64 *tm = s_mask + ew * (top - 1);
            for (size_t w = 0; w < ew; w++) tm[w] |= carry[w];
            if (carry_allseq) s_allchild[top - 1] = 1;
            /* node complete: [s_lb, i-1], depth s_lcp */
            int allseq = 1;
            for (int w = 0; w < W; w++) if (tm[w] != full[w]) allseq = 0;
            if (allseq && !s_allchild[top - 1]) {
                int sfx = 0;
                for (int c = 0; c < 5 && !sfx; c++) {
                    int all = 1;
                    for (int w = 0; w < W; w++) if (tm[(size_t)(1 + c) * W + w] != full[w]) all = 0;
                    if (all) sfx = 1;
                }
                if (o->nnodes == o->ncap) {
                    o->ncap = o->ncap ? o->ncap * 2 : 256;
                    o->nodes = (cnode *)realloc(o->nodes, sizeof(cnode) * o->ncap);
                }
                cnode *c = &o->nodes[o->nnodes++];
                c->lb = s_lb[top - 1]; c->rb = i - 1; c->depth = s_lcp[top - 1]; c->dfs = 0; c->gone = 0; c->sfx = sfx;
            }
            memcpy(carry, tm, sizeof(u64) * ew);
            carry_allseq = allseq;
            lb = s_lb[top - 1];
            top--;
        }
        if (top == 0) break; /* root popped (i==N) */
        if (cur > s_lcp[top - 1]) {
            if (top == cap) {
                cap *= 2;
                s_lcp = (int *)realloc(s_lcp, sizeof(int) * cap);
                s_lb = (int *)realloc(s_lb, sizeof(int) * cap);
                s_allchild = (char *)realloc(s_allchild, cap);
                s_mask = (u64 *)realloc(s_mask, sizeof(u64) * ew * cap);
            }
            s_lcp[top] = cur; s_lb[top] = lb; s_allchild[top] = (char)carry_allseq;
            memcpy(s_mask + ew * top, carry, sizeof(u64) * ew);
            top++;
        } else {
            u64 *tm = s_mask + ew * (top - 1);
            for (size_t w = 0; w < ew; w++) tm[w] |= carry[w];
            if (carry_allseq) s_allchild[top - 1] = 1;
        }
    }
    o->count_collected = o->nnodes;
    free(s_lcp); free(s_lb); free(s_allchild); free(s_mask); free(carry); free(full);
}

static int cmp_cnode(const void *a, const void *b) {
    const cnode *x = (const cnode *)a, *y = (const cnode *)b;
    if (x->depth != y->depth) return (x->depth > y->depth) ? -1 : 1; /* nodeslinkedlists.c:36 insertSortedItem */
    if (x->dfs != y->dfs) return (x->dfs > y->dfs) ? -1 : 1;         /* later visited goes first */
    return 0;
}

/* csamsa.c:80 removeSuffixNodes.  `nodes` is blockslist in list order.
 *
 * What the reference does: for every list item Y in turn it follows Y[1:], Y[2:], ... (getSuffixNode,
 * gencycsuffixtrees.c:327: the node AT OR BELOW the place where the string ends) and deletes a later
 * list item when it IS that node; one pointer (searchnode) sweeps the list, forward only, for all
 * suffixes of one Y.
 *
 * When no list item is a leaf (no whole rotation occurs in every sequence) every Y is a branching
 * node, so is each Y[j:], the sweep never misses, and the outcome is: X goes iff x.X occurs in all
 * sequences for some letter x (`sfx`, from the scan) -- that is the rule the CUDA path uses, and
 * CSA_ORACLE_LITERAL=1 makes this function replay the list walk instead, to check the two against
 * each other.  With leaves on the list the place where Y[j:] ends can lie inside the edge to the
 * leaf of the NEXT rotation, which is then deleted although it is no suffix of Y: those sets are
 * always replayed literally. */
static void remove_suffix_nodes(const gtext *t, const int *sa, const int *lcp, const int *isa, int N, scan_out *o) {
    int C = o->nnodes, literal = getenv("CSA_ORACLE_LITERAL") != NULL;
    cnode *c = o->nodes;
    o->count_suffixfree = C;
    if (C == 0 || c[0].depth == 0) return; /* csamsa.c:85: a list that holds only the root is left alone */
    if (c[0].depth >= t->nmin) literal = 1;
    if (!literal) {
        for (int i = 0; i < C; i++) if (c[i].sfx) { c[i].gone = 1; o->count_suffixfree--; }
        return;
    }
    int *nxt = (int *)malloc(sizeof(int) * C), *prv = (int *)malloc(sizeof(int) * C);
    int *at = (int *)malloc(sizeof(int) * (size_t)(N + 1));
    for (int i = 0; i <= N; i++) at[i] = -1;
    for (int i = 0; i < C; i++) { nxt[i] = i + 1 < C ? i + 1 : -1; prv[i] = i - 1; at[c[i].lb] = i; }
    for (int node = 0; node != -1; node = nxt[node]) {
        int g = sa[c[node].lb], len = c[node].depth;
        int search = nxt[node];
        for (;;) {
            /* getSuffixNode: the string loses its first letter */
            { int k = t->seqof[g]; g = t->off[k] + (g - t->off[k] + 1) % t->per[k]; } /* (w^c: the leaf of rotation p mod |w|) */
            len--;
            if (search == -1 || len <= 0) break; /* csamsa.c:92 `searchnode!=NULL && suffix!=root` */
            int l = isa[g], r = l, d;
            while (l > 0 && lcp[l] >= len) l--;
            while (r + 1 < N && lcp[r + 1] >= len) r++;
            if (l == r) d = t->n[t->seqof[sa[l]]];
            else { d = INT_MAX; for (int q = l + 1; q <= r; q++) if (lcp[q] < d) d = lcp[q]; }
            while (search != -1 && d < c[search].depth) search = nxt[search];
            while (search != -1 && d == c[search].depth) {
                if (c[search].lb == l && c[search].rb == r) {
                    int del = search, p = prv[del], nx = nxt[del];
                    if (nx != -1) prv[nx] = p;
                    if (p != -1) nxt[p] = nx;
                    search = p != -1 ? p : nx; /* nodeslinkedlists.c:88 deleteItem returns prev, else next */
                    c[del].gone = 1;
                    o->count_suffixfree--;
                    if (del == node) o->undefined = 1; /* the reference goes on reading the freed item */
                    break;
                }
                search = nxt[search];
            }
        }
        if (o->undefined) break;
    }
    if (getenv("CSA_ORACLE_TRACE")) {
        int dl = 0, dn = 0, nl = 0;
        for (int i = 0; i < C; i++) {
            if (c[i].depth >= t->nmin) { nl++; if (c[i].gone) dl++; }
            else if (c[i].gone != c[i].sfx) dn++;
        }
        fprintf(stderr, "[oracle] literal list walk: %d leaves on the list, %d of them deleted; %d inner nodes differ from the simple rule; undefined=%d\n", nl, dl, dn, o->undefined);
    }
    free(nxt); free(prv); free(at);
}

/* csamsa.c:283 removeNonUniqueNodes: exactly one leaf per sequence below the node */
static void keep_unique_nodes(const gtext *t, const int *sa, scan_out *o) {
    int m = t->m;
    for (int i = 0; i < o->nnodes; i++) {
        cnode *c = &o->nodes[i];
        if (c->gone || c->rb - c->lb + 1 != m) continue;
        if (o->nblocks == o->cap) {
            o->cap = o->cap ? o->cap * 2 : 256;
            o->blocks = (blk *)realloc(o->blocks, sizeof(blk) * o->cap);
        }
        blk *b = &o->blocks[o->nblocks++];
        b->depth = c->depth; b->lb = c->lb; b->dfs = c->dfs;
        b->pos = (int *)malloc(sizeof(int) * m);
        for (int j = c->lb; j <= c->rb; j++) {
            int gg = sa[j], kk = t->seqof[gg];
            b->pos[kk] = gg - t->off[kk];
        }
        o->count_unique++;
    }
}

/* ---- DFS order of the reference's tree, restricted to what decides block order -------
 * Children of a node are stored in creation order (gencycsuffixtrees.c:193 addBranch appends,
 * :210 splitNode keeps the slot), and Ukkonen's phase ii creates the branch for string S.c when
 * the first occurrence of S.c ends at ii.  Sequence 0 is inserted first and every ancestor of
 * a block contains sequence 0, so two blocks are visited in the order of the first occurrence,
 * in sequence 0, of the strings on which they diverge.  dfs_rank_seq0 numbers all rotations of
 * sequence 0 in that visiting order. */
static void dfs_rank_seq0(const gtext *t, const int *sa, const int *lcp, int N, int *dfsrank /* n0 */) {
    int n0 = t->per[0]; /* leaves of sequence 0 */
    int *sa0 = (int *)malloc(sizeof(int) * n0), *lcp0 = (int *)malloc(sizeof(int) * n0);
    int c = 0, run = INT_MAX;
    for (int i = 0; i < N; i++) {
        if (i > 0 && lcp[i] < run) run = lcp[i];
        if (t->seqof[sa[i]] == 0) {
            sa0[c] = sa[i] - t->off[0];
            lcp0[c] = (c == 0) ? 0 : run;
            c++;
            run = INT_MAX;
        }
    }
    /* explicit stack of [l,r] intervals */
    int cap = 1024, top = 0;
    int *st = (int *)malloc(sizeof(int) * 2 * cap);
    int counter = 0;
    st[0] = 0; st[1] = n0 - 1; top = 1;
    int ccap = 64;
    int *cl = (int *)malloc(sizeof(int) * ccap), *cr = (int *)malloc(sizeof(int) * ccap),
        *cm = (int *)malloc(sizeof(int) * ccap);
    while (top > 0) {
        top--;
        int l = st[2 * top], r = st[2 * top + 1];
        if (l == r) { dfsrank[sa0[l]] = counter++; continue; }
        int mn = INT_MAX;
        for (int i = l + 1; i <= r; i++) if (lcp0[i] < mn) mn = lcp0[i];
        int nc = 0, start = l;
        for (int i = l + 1; i <= r + 1; i++) {
            if (i == r + 1 || lcp0[i] == mn) {
                if (nc == ccap) {
                    ccap *= 2;
                    cl = (int *)realloc(cl, sizeof(int) * ccap);
                    cr = (int *)realloc(cr, sizeof(int) * ccap);
                    cm = (int *)realloc(cm, sizeof(int) * ccap);
                }
                int mp = INT_MAX;
                for (int j = start; j < i; j++) if (sa0[j] < mp) mp = sa0[j];
                cl[nc] = start; cr[nc] = i - 1; cm[nc] = mp; nc++;
                start = i;
            }
        }
        /* insertion sort children by first occurrence, then push in reverse */
        for (int a = 1; a < nc; a++) {
            int tl = cl[a], tr = cr[a], tm = cm[a], b = a - 1;
            while (b >= 0 && cm[b] > tm) { cl[b + 1] = cl[b]; cr[b + 1] = cr[b]; cm[b + 1] = cm[b]; b--; }
            cl[b + 1] = tl; cr[b + 1] = tr; cm[b + 1] = tm;
        }
        while (top + nc > cap) { cap *= 2; st = (int *)realloc(st, sizeof(int) * 2 * cap); }
        for (int a = nc - 1; a >= 0; a--) { st[2 * top] = cl[a]; st[2 * top + 1] = cr[a]; top++; }
    }
    free(sa0); free(lcp0); free(st); free(cl); free(cr); free(cm);
}

/* ---- collectNodeChains (csamsa.c:135-279), literal ---------------------------------- */
typedef struct ekey { int e, b; } ekey;
static int cmp_ekey(const void *a, const void *b) {
    const ekey *x = (const ekey *)a, *y = (const ekey *)b;
    if (x->e != y->e) return x->e < y->e ? -1 : 1;
    return x->b - y->b;
}

static int chain_blocks(const gtext *t, scan_out *o, const char *bad, int max_interval, csa_oracle_result *res) {
    int B = o->nblocks, m = t->m;
    blk *bl = o->blocks;
    int *size = (int *)calloc(B, sizeof(int)), *total = (int *)calloc(B, sizeof(int));
    int *interval = (int *)calloc(B, sizeof(int)), *next = (int *)malloc(sizeof(int) * B);
    int mcs = B;
    int hang = 0;
    for (int b = 0; b < B; b++) next[b] = -1;
    /* csamsa.c:147-183: walk the text of every sequence (twice round at most) through the tree; a block
     * is noticed at the text index that follows an occurrence of it (e = start+depth, unrolled
     * coordinates); the walk stops at textsize + start of the first block noticed (:168 n+=...).
     * A sequence w^c meets each of its blocks c times a lap (one leaf, c places in the text).
     * bad[]: places where the walk stands on a LEAF that holds every sequence (a whole rotation of the
     * shortest sequence inside all others).  From there csamsa.c:176 takes the leaf's suffixlink -- which
     * for a leaf is the link to the NEXT ROTATION's leaf (gencycsuffixtrees.c:505), depth unchanged -- and
     * retries the same letter, leaf after leaf, until a leaf has no such branch (followChar NULL ->
     * nodeFromAllSeqs(NULL), csamsa.c:153: SIGSEGV) or for ever.  Either way the reference never
     * completes: -2. */
    for (int k = 0; k < m; k++) {
        int n = t->n[k], per = t->per[k], reps = 2 * (n / per);
        size_t ne = 0, cap = (size_t)B * reps + 2 * (size_t)n + 1;
        ekey *ek = (ekey *)malloc(sizeof(ekey) * cap);
        for (int b = 0; b < B; b++)
            for (int r = 0; r < reps; r++) { ek[ne].e = bl[b].pos[k] + r * per + bl[b].depth; ek[ne].b = b; ne++; }
        if (bad)
            for (int q = 0; q < 2 * n; q++)
                if (bad[t->off[k] + q % per]) { ek[ne].e = q + t->nmin; ek[ne].b = -1; ne++; }
        qsort(ek, ne, sizeof(ekey), cmp_ekey);
        int limit = n, first = 1;
        int prev = -1;
        for (size_t j = 0; j < ne && ek[j].e < limit; j++) {
            int b = ek[j].b;
            if (b < 0) { hang = -2; break; }
            if (first) { limit = n + (ek[j].e - bl[b].depth); first = 0; }
            if (prev != -1 && size[prev] == 0) {
                if (next[prev] == -1) next[prev] = b;
                else if (next[prev] != b) { next[prev] = -1; size[prev] = -1; }
            }
            prev = b;
        }
        free(ek);
        if (hang) break;
    }
    /* csamsa.c:185-233 */
    for (int b = 0; b < B && !hang; b++) {
        if (total[b] == -1) continue;
        size[b] = bl[b].depth;
        int prev = b, cur = next[b];
        long long guard = 0;
        while (cur != -1) {
            if (++guard > 4LL * B + 16) { hang = 1; break; }
            int iv = INT_MAX;
            for (int k = 0; k < m; k++) {
                int count = 0;
                if (bl[cur].pos[k] < bl[prev].pos[k]) count += t->n[k];
                count += bl[cur].pos[k] - (bl[prev].pos[k] + bl[prev].depth);
                if (count < iv) iv = count;
            }
            if (iv > max_interval) { next[prev] = -1; break; }
            if (total[cur] > 0) {
                size[b] += size[cur];
                total[b] += total[cur];
                interval[prev] = iv;
                total[b] += iv;
                size[cur] = bl[cur].depth;
                total[cur] = -1;
                mcs--;
                break;
            }
            size[cur] = bl[cur].depth;
            size[b] += size[cur];
            interval[prev] = iv;
            total[b] += iv;
            total[cur] = -1;
            mcs--;
            prev = cur;
            cur = next[cur];
        }
        total[b] += size[b];
    }
    /* nodeslinkedlists.c:59 sortList: repeatedly move the first strictly-largest to the front
     * == stable sort by size, descending */
    int *ord = (int *)malloc(sizeof(int) * (B ? B : 1)), *inv = (int *)malloc(sizeof(int) * (B ? B : 1));
    for (int b = 0; b < B; b++) ord[b] = b;
    for (int a = 1; a < B; a++) { /* insertion sort keeps ties in list order */
        int v = ord[a], j = a - 1;
        while (j >= 0 && size[ord[j]] < size[v]) { ord[j + 1] = ord[j]; j--; }
        ord[j + 1] = v;
    }
    for (int i = 0; i < B; i++) inv[ord[i]] = i;
    res->nblocks = B;
    res->count_chains = mcs;
    res->depth = (int *)malloc(sizeof(int) * (B ? B : 1));
    res->size = (int *)malloc(sizeof(int) * (B ? B : 1));
    res->totalsize = (int *)malloc(sizeof(int) * (B ? B : 1));
    res->interval = (int *)malloc(sizeof(int) * (B ? B : 1));
    res->next = (int *)malloc(sizeof(int) * (B ? B : 1));
    res->positions = (int *)malloc(sizeof(int) * (size_t)(B ? B : 1) * m);
    for (int i = 0; i < B; i++) {
        int b = ord[i];
        res->depth[i] = bl[b].depth;
        res->size[i] = size[b];
        res->totalsize[i] = total[b];
        res->interval[i] = interval[b];
        res->next[i] = next[b] == -1 ? -1 : inv[next[b]];
        memcpy(res->positions + (size_t)i * m, bl[b].pos, sizeof(int) * m);
    }
    free(size); free(total); free(interval); free(next); free(ord); free(inv);
    return hang;
}

/* identical rotations of one sequence share a leaf (see gtext_init): drop (k,p) for p >= per[k].  The
 * dropped entries stand right behind their representative (equal strings, ties by index), so the
 * LCP of the next kept entry is the minimum over the dropped run. */
static int collapse_periodic(const gtext *t, int *sa, int *lcp) {
    int N = t->N, w = 0, run = INT_MAX;
    for (int i = 0; i < N; i++) {
        int k = t->seqof[sa[i]];
        if (lcp[i] < run) run = lcp[i];
        if (sa[i] - t->off[k] >= t->per[k]) continue;
        sa[w] = sa[i];
        lcp[w] = (w == 0) ? 0 : run;
        w++;
        run = INT_MAX;
    }
    return w;
}

/* places (k,p) whose first nmin letters are a whole rotation of the shortest sequence AND occur in every
 * sequence: the leaves of the tree that hold all sequences (see chain_blocks) */
static char *leaves_of_all(const gtext *t, const int *sa, const int *lcp, int N) {
    char *bad = (char *)calloc(t->N ? t->N : 1, 1);
    int any = 0, W = (t->m + 63) / 64;
    u64 *seen = (u64 *)malloc(sizeof(u64) * W);
    for (int i = 0; i < N;) {
        int j = i + 1;
        while (j < N && lcp[j] >= t->nmin) j++;
        if (j - i >= t->m) {
            int cnt = 0;
            memset(seen, 0, sizeof(u64) * W);
            for (int q = i; q < j; q++) {
                int k = t->seqof[sa[q]];
                if (!(seen[k >> 6] >> (k & 63) & 1)) { seen[k >> 6] |= 1ULL << (k & 63); cnt++; }
            }
            if (cnt == t->m) { for (int q = i; q < j; q++) bad[sa[q]] = 1; any = 1; }
        }
        i = j;
    }
    free(seen);
    if (!any) { free(bad); return NULL; }
    return bad;
}

/* nodeslinkedlists.c:144-165 blockLabel spells a block from the labels of the tree edges on its path,
 * i.e. from the text that CREATED each edge (labelfrom/startpos, gencycsuffixtrees.c:160-218).  The
 * edge that holds letter number l of block X was created by the first rotation, in insertion order,
 * that begins with X[0..l]: sequence 0 (every block occurs there), at the smallest such position p:
 * letter = texts[0][p+l].  For A/C/G/T every occurrence holds the same letter; a letter outside
 * ACGT is spelled as that first occurrence has it. */
static char *block_letters(const gtext *t, const char *const *texts, const int *sa, const int *lcp, int N,
                           const int *isa_c, const blk *b) {
    int d = b->depth, p0 = b->pos[0], n0 = t->n[0];
    char *out = (char *)malloc((size_t)d + 1);
    int l = isa_c[t->off[0] + p0], r = l, minpos = p0;
    for (int j = d - 1; j >= 0; j--) {
        char c = texts[0][(p0 + j) % n0];
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') { out[j] = c; continue; }
        while (l > 0 && lcp[l] >= j + 1) {
            l--;
            if (t->seqof[sa[l]] == 0 && sa[l] - t->off[0] < minpos) minpos = sa[l] - t->off[0];
        }
        while (r + 1 < N && lcp[r + 1] >= j + 1) {
            r++;
            if (t->seqof[sa[r]] == 0 && sa[r] - t->off[0] < minpos) minpos = sa[r] - t->off[0];
        }
        out[j] = texts[0][(minpos + j) % n0];
    }
    out[d] = 0;
    return out;
}

int csa_oracle_run(int m, const char *const *texts, const int *textsizes, int max_interval,
                   csa_oracle_result *res) {
    return csa_oracle_run_sa(m, texts, textsizes, max_interval, res, NULL, NULL);
}

/* the same, and the suffix array + LCP array it worked on (layout of csa_oracle_gsa) copied out */
int csa_oracle_run_sa(int m, const char *const *texts, const int *textsizes, int max_interval,
                      csa_oracle_result *res, int *sa_out, int *lcp_out) {
    gtext t;
    memset(res, 0, sizeof(*res));
    res->m = m;
    if (m < 2 || gtext_init(&t, m, texts, textsizes) != 0) { res->status = -1; return -1; }
    int N = t.N;
    int *sa = (int *)malloc(sizeof(int) * N), *isa = (int *)malloc(sizeof(int) * N);
    int *lcp = (int *)malloc(sizeof(int) * N);
    build_gsa(&t, sa, isa);
    build_lcp(&t, sa, isa, lcp);
    if (sa_out) { /* dropped rotations behind all others, lcp 0 */
        int nd = 0, *dups = (int *)malloc(sizeof(int) * (N ? N : 1));
        for (int i = 0; i < N; i++) {
            int k = t.seqof[sa[i]];
            if (sa[i] - t.off[k] >= t.per[k]) dups[nd++] = sa[i];
        }
        int w = collapse_periodic(&t, sa, lcp);
        memcpy(sa_out, sa, sizeof(int) * w); memcpy(lcp_out, lcp, sizeof(int) * w);
        for (int j = 0; j < nd; j++) { sa_out[w + j] = dups[j]; lcp_out[w + j] = 0; }
        free(dups);
        N = w;
    } else
        N = collapse_periodic(&t, sa, lcp);
    for (int i = 0; i < N; i++) isa[sa[i]] = i;
    scan_out o;
    scan_intervals(&t, sa, lcp, N, &o);
    {   /* list order: depth, then the DFS of csamsa.c:64 (any rotation of sequence 0 below the node numbers it) */
        int *dfsrank = (int *)malloc(sizeof(int) * t.n[0]);
        dfs_rank_seq0(&t, sa, lcp, N, dfsrank);
        for (int i = 0; i < o.nnodes; i++) {
            int q = o.nodes[i].lb;
            while (t.seqof[sa[q]] != 0) q++;
            o.nodes[i].dfs = dfsrank[sa[q] - t.off[0]];
        }
        free(dfsrank);
        qsort(o.nodes, o.nnodes, sizeof(cnode), cmp_cnode);
    }
    remove_suffix_nodes(&t, sa, lcp, isa, N, &o);
    keep_unique_nodes(&t, sa, &o);
    res->count_collected = o.count_collected;
    res->count_suffixfree = o.count_suffixfree;
    res->count_unique = o.count_unique;
    if (o.count_collected == 0) res->status = CSA_ORACLE_NO_COMMON;
    else if (o.undefined) res->status = CSA_ORACLE_UNDEFINED;
    else if (o.count_unique == 0) res->status = CSA_ORACLE_NO_UNIQUE;
    if (res->status == 0) {
        char *bad = leaves_of_all(&t, sa, lcp, N);
        int rc = chain_blocks(&t, &o, bad, max_interval, res);
        free(bad);
        if (rc == -2) res->status = CSA_ORACLE_DEGENERATE;
        else if (rc) res->status = CSA_ORACLE_HANG;
        else {
            /* csamsa.c:311 getRotations: positions of the head of the sorted list */
            res->rotations = (int *)malloc(sizeof(int) * m);
            for (int k = 0; k < m; k++) res->rotations[k] = res->positions[k];
            res->letters = (char **)calloc(res->nblocks ? res->nblocks : 1, sizeof(char *));
            for (int i = 0; i < res->nblocks; i++) {
                blk tmp;
                tmp.depth = res->depth[i];
                tmp.pos = res->positions + (size_t)i * m;
                res->letters[i] = block_letters(&t, texts, sa, lcp, N, isa, &tmp);
            }
        }
    }
    for (int b = 0; b < o.nblocks; b++) free(o.blocks[b].pos);
    free(o.blocks);
    free(o.nodes);
    free(sa); free(isa); free(lcp);
    gtext_free(&t);
    return res->status;
}

void csa_oracle_free(csa_oracle_result *r) {
    free(r->depth); free(r->size); free(r->totalsize); free(r->interval); free(r->next);
    free(r->positions); free(r->rotations);
    if (r->letters) { for (int i = 0; i < r->nblocks; i++) free(r->letters[i]); free(r->letters); }
    memset(r, 0, sizeof(*r));
}

/* does the chain that starts at block b close into a ring?  blockLabel (nodeslinkedlists.c:150
 * `while(tmpblock!=NULL)`) then never stops writing: the reference dies there, after -Rotated.fasta */
int csa_oracle_chain_is_ring(const csa_oracle_result *r, int b) {
    int guard = 0;
    for (int cur = b; cur != -1; cur = r->next[cur])
        if (++guard > r->nblocks) return 1;
    return 0;
}

/* nodeslinkedlists.c:128 blockLabel: the blocks of the chain, the gaps between them as dashes */
char *csa_oracle_block_label(const csa_oracle_result *r, int b) {
    size_t cap = 256, len = 0; /* len plays labelpos */
    char *label = (char *)calloc(cap, 1);
    int guard = 0;
    for (int cur = b; cur != -1 && guard <= r->nblocks; cur = r->next[cur], guard++) {
        int d = r->depth[cur];
        while (len + d + 32 > cap) {
            label = (char *)realloc(label, cap * 2);
            memset(label + cap, 0, cap);
            cap *= 2;
        }
        memcpy(label + len, r->letters[cur], (size_t)d);
        len += d;
        int n = r->interval[cur];
        if (n < 0) {
            len = ((long long)len + n < 0) ? 0 : len + n;
        } else if (n > 7) {
            len += (size_t)sprintf(label + len, "-(%d)-", n);
        } else {
            for (int i = 0; i < n; i++) label[len++] = '-';
        }
    }
    label[len] = '\0';
    return label;
}

/* The suffix array the block stages work on, in the layout the CUDA path keeps (csa_gpu_batch_suffix_array): the
 * identical rotations of a sequence w^c (p >= |w|, collapse_periodic above) stand behind all others, lcp 0. */
int csa_oracle_gsa(int m, const char *const *texts, const int *textsizes, int *sa_out,
                   int *lcp_out) {
    gtext t;
    if (gtext_init(&t, m, texts, textsizes) != 0) return -1;
    int N = t.N, nd = 0;
    int *isa = (int *)malloc(sizeof(int) * N), *dups = (int *)malloc(sizeof(int) * N);
    build_gsa(&t, sa_out, isa);
    build_lcp(&t, sa_out, isa, lcp_out);
    for (int i = 0; i < N; i++) {
        int k = t.seqof[sa_out[i]];
        if (sa_out[i] - t.off[k] >= t.per[k]) dups[nd++] = sa_out[i];
    }
    int w = collapse_periodic(&t, sa_out, lcp_out);
    for (int j = 0; j < nd; j++) { sa_out[w + j] = dups[j]; lcp_out[w + j] = 0; }
    free(isa); free(dups);
    gtext_free(&t);
    return 0;
}
