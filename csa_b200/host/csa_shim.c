/*
 * csa_shim.c -- the reference-side binding, COMPILED: one translation unit that, linked with the
 * reference's own unmodified objects and libcsa_gpu.so, makes `./CSA R <multi-fasta>` run its hot path
 * on the B200.  It replaces exactly the two calls of csamsa.c:599 and :610
 *
 *     tree = buildGeneralizedTree();   gencycsuffixtrees.c:418
 *     analyzeTree();                   csamsa.c:324
 *
 * and hands back what they leave behind in the reference's own globals -- `rotations` (csamsa.h:12),
 * the sorted `blockslist` (csamsa.c:30) of linkedblock items (nodeslinkedlists.h:4-13) and `mcscount` --
 * so that everything downstream is the REFERENCE's code, untouched: saveRotatedSequences (csamsa.c:421)
 * and createImageAndShowResults (csamsa.c:361) with blockLabel, initializeBlocks, drawBlockRotated,
 * finalizeGraphics: <base>-Rotated.fasta, -Blocks.csv, -positions.txt, -imagemap.txt, -Blocks.bmp.
 *
 * Built (recipe in INTEGRATION.md) where the reference sources are present: their objects are
 * compiled from /root/reference/source/*.c as they lie there, the two functions above are made weak in
 * those objects (objcopy --weaken-symbol; no source line is touched) and the strong definitions below
 * win at link time.  A maintainer of the reference would instead delete the two bodies and add this
 * file plus -lcsa_gpu to source/Makefile.
 *
 * This file includes the reference's headers (for its struct layouts and globals); it holds none of
 * its code.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "csamsa.h"            /* -I<reference>/source */
#include "gencycsuffixtrees.h"
#include "nodeslinkedlists.h"
#include "csa_gpu.h"           /* -I<repo>/include */

extern struct _linkedblock *blockslist; /* csamsa.c:30 */
extern int mcscount;                    /* csamsa.c:33 */

static csa_gpu_ctx *g_ctx;
static csa_gpu_set_info g_info;
static int *g_rot;

static void stop(int code, const char *why) { /* where the reference itself would die or hang */
    printf("\n> CSA_GPU: %s\n", why);
    fflush(stdout);
    exit(code);
}

static void fail_gpu(const char *what) {
    fprintf(stderr, "> ERROR: %s: %s (this build has no CPU path)\n", what, csa_gpu_last_error());
    exit(70);
}

static int letter_class(int c) { return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : '-'; }

/* is b a rotation of a (same length), letters outside ACGT all alike (gencycsuffixtrees.c:283)? */
static int is_rotation(const char *a, const char *b, int n) {
    char *dbl = (char *)malloc(2 * (size_t)n + 1), *pat = (char *)malloc((size_t)n + 1);
    int q, hit;
    for (q = 0; q < 2 * n; q++) dbl[q] = (char)letter_class(a[q % n]);
    for (q = 0; q < n; q++) pat[q] = (char)letter_class(b[q]);
    dbl[2 * n] = pat[n] = 0;
    hit = strstr(dbl, pat) != NULL;
    free(dbl);
    free(pat);
    return hit;
}

/* gencycsuffixtrees.c:418.  Of the construction only its bookkeeping is left: a sequence that is an
 * identical rotation of an earlier one is dropped (:518-524, same warning, same progress dots), the
 * device does the rest in analyzeTree.  Returns a root that freeTreeNode (csamsa.c:655) can free. */
treenode *buildGeneralizedTree() {
    int i, j, *shown = (int *)malloc(sizeof(int) * (size_t)numberofseqs);
    for (i = 0; i < numberofseqs; i++) shown[i] = i;
    for (j = 0; j < numberofseqs; j++) {
        int dup = -1;
        for (i = 0; i < j && dup < 0; i++)
            if (textsizes[i] == textsizes[j] && is_rotation(texts[i], texts[j], textsizes[j])) dup = i;
        if (dup >= 0) {
            printf("> WARNING: Discarding seq. %d because it is an identical rotation of seq. %d\n", shown[j] + 1, shown[dup] + 1);
            for (i = j; i < numberofseqs; i++) shown[i]++;
            free(texts[j]);
            free(descs[j]);
            for (i = j + 1; i < numberofseqs; i++) {
                texts[i - 1] = texts[i]; descs[i - 1] = descs[i]; textsizes[i - 1] = textsizes[i];
            }
            numberofseqs--;
            if (numberofseqs < 2) exitMessage("The program needs at least 2 sequences to run");
            j--;
        }
        printf(".");
        fflush(stdout);
    }
    free(shown);
    root = (treenode *)calloc(1, sizeof(treenode));
    root->startpos = root->endpos = -1;
    root->leaffrom = -1;
    return root;
}

/* csamsa.c:324: the same four progress lines with the same counts, then the globals filled */
void analyzeTree() {
    int device = getenv("CSA_GPU_DEVICE") ? atoi(getenv("CSA_GPU_DEVICE")) : 0;
    int m = numberofseqs, nb, b, k;
    int *depth, *size, *total, *interval, *next, *positions;
    long long nletters, *off;
    char *letters;
    linkedblock **item;
    blockslist = NULL;
    mcscount = 0;
    printf("> Collecting maximum common subsequences... ");
    fflush(stdout);
    if (csa_gpu_create(device, &g_ctx) != CSA_GPU_OK) fail_gpu("no CUDA device");
    g_rot = (int *)calloc((size_t)m, sizeof(int));
    if (csa_gpu_find_rotations(g_ctx, m, (const char *const *)texts, textsizes, INT_MAX, CSA_GPU_FLAG_STATS, g_rot, &g_info) != CSA_GPU_OK)
        fail_gpu("GPU run failed");
    if (g_info.count_collected == 0) exitMessage("No common subsequences found");
    printf("%d nodes found\n", g_info.count_collected);
    printf("> Removing suffixes... ");
    fflush(stdout);
    if (g_info.status == CSA_SET_UNDEFINED)
        stop(6, "removeSuffixNodes (csamsa.c:80) frees the list item it stands on here; the reference's answer is not defined");
    printf("%d nodes left\n", g_info.count_suffixfree);
    printf("> Removing repeats... ");
    fflush(stdout);
    if (g_info.count_unique == 0) exitMessage("No unique subsequences found");
    printf("%d nodes left\n", g_info.count_unique);
    printf("> Connecting block chains... ");
    fflush(stdout);
    if (g_info.status == CSA_SET_DEGENERATE)
        stop(3, "collectNodeChains walks off a leaf here (csamsa.c:153): a whole rotation of the shortest sequence occurs in all others");
    if (g_info.status == CSA_SET_NONTERMINATING)
        stop(4, "collectNodeChains (csamsa.c:197) does not terminate on this input: the common blocks form a cycle");
    printf("%d chains found\n", g_info.count_chains);
    mcscount = g_info.count_chains;

    /* ---- blockslist: one linkedblock per block, in the order of the reference's sorted list ---- */
    nb = g_info.nblocks;
    depth = (int *)calloc((size_t)nb, sizeof(int)); size = (int *)calloc((size_t)nb, sizeof(int));
    total = (int *)calloc((size_t)nb, sizeof(int)); interval = (int *)calloc((size_t)nb, sizeof(int));
    next = (int *)calloc((size_t)nb, sizeof(int)); positions = (int *)calloc((size_t)nb * (size_t)m, sizeof(int));
    off = (long long *)calloc((size_t)nb + 1, sizeof(long long));
    if (csa_gpu_batch_blocks(g_ctx, depth, size, total, interval, next, positions) != CSA_GPU_OK) fail_gpu("GPU run failed");
    nletters = csa_gpu_batch_block_letters(g_ctx, NULL, off);
    if (nletters < 0) fail_gpu("GPU run failed");
    letters = (char *)calloc((size_t)nletters + 1, 1);
    if (nletters && csa_gpu_batch_block_letters(g_ctx, letters, NULL) < 0) fail_gpu("GPU run failed");
    /* blockLabel (nodeslinkedlists.c:154-165) reads a block's letters from texts[item->labelfrom] between
     * item->startpos and item->endpos, up the backlinks to the root: every block gets one node hanging off
     * the root whose label lies in one extra text -- the letters exactly as the reference's tree would spell
     * them (csa_gpu_batch_block_letters) -- stored behind the sequences (not counted in numberofseqs). */
    texts = (char **)realloc(texts, sizeof(char *) * (size_t)(m + 1));
    textsizes = (int *)realloc(textsizes, sizeof(int) * (size_t)(m + 1));
    texts[m] = letters;
    textsizes[m] = (int)nletters + 1;
    item = (linkedblock **)calloc((size_t)nb + 1, sizeof(linkedblock *));
    for (b = 0; b < nb; b++) {
        treenode *node = (treenode *)calloc(1, sizeof(treenode));
        node->id = b + 1;
        node->depth = depth[b];
        node->startpos = (int)off[b];
        node->endpos = (int)off[b + 1] - 1;
        node->labelfrom = m;
        node->leaffrom = -1;
        node->backlink = root;
        item[b] = createItem(node); /* nodeslinkedlists.c:10 */
        item[b]->size = size[b];
        item[b]->totalsize = total[b];
        item[b]->interval = interval[b];
        item[b]->positions = (int *)calloc((size_t)m, sizeof(int));
        for (k = 0; k < m; k++) item[b]->positions[k] = positions[(size_t)b * m + k];
        node->fromblock = item[b];
    }
    for (b = 0; b < nb; b++) {
        item[b]->prev = b ? item[b - 1] : NULL;
        item[b]->next = b + 1 < nb ? item[b + 1] : NULL;
        item[b]->nextblock = next[b] >= 0 ? item[next[b]] : NULL;
    }
    blockslist = nb ? item[0] : NULL;
    /* csamsa.c:311 getRotations */
    rotations = g_rot;
    free(depth); free(size); free(total); free(interval); free(next); free(positions); free(off); free(item);
    csa_gpu_destroy(g_ctx);
    g_ctx = NULL;
}
