/*
 * csa_main.c -- `CSA R <multi-fasta-file>`: the reference's rotation-only command line on the B200.
 *
 * Host code in C, as in the reference; everything between loading and writing is ONE call into
 * libcsa_gpu.so (include/csa_gpu.h), which stands where the reference calls
 * buildGeneralizedTree() and analyzeTree() (csamsa.c:599,610).  What this file mirrors:
 *   csamsa.c:437  LoadSequences          same parsing, same lines on stdout
 *   gencycsuffixtrees.c:518-524          a sequence that is a rotation of an earlier one is dropped
 *   csamsa.c:324  analyzeTree            the four progress lines with their counts
 *   csamsa.c:421  saveRotatedSequences   <base>-Rotated.fasta, byte for byte
 *   csamsa.c:361  createImageAndShowResults: <base>-Blocks.csv and the chain list on stdout
 * Not produced HERE: the .bmp picture and its -positions / -imagemap side files (graphics.c, bitmap.c:
 * drawing code) and the other modes (A, I, C, S, M).  csa_shim.c is the other way to the same library:
 * linked with the reference's own objects it keeps the reference's main, loader and drawing code and
 * writes all five files (INTEGRATION.md has the build recipe).
 * There is no CPU fallback: without a CUDA device the program stops with an error (exit 70).
 * Exit codes: 0 as the reference (its error messages included, csamsa.c:57 exits 0); 3, 4, 5, 6 where the
 * reference itself would crash or never return (see CSA_SET_* in csa_gpu.h); 70 accelerator failure.
 */
#include "../../include/csa_gpu.h"
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXNUMBEROFSEQS 64 /* csamsa.c:22 */

static const char *inputfilename;

static void die(const char *msg) { /* csamsa.c:57 exitMessage */
    printf("\n> ERROR: %s\n", msg);
    exit(0);
}

static void stop(int code, const char *why) { /* where the reference itself would die or hang */
    printf("\n> CSA_GPU: %s\n", why);
    fflush(stdout);
    exit(code);
}

static void fail_gpu(const char *what) {
    fflush(stdout);
    fprintf(stderr, "> ERROR: %s: %s (this build has no CPU path)\n", what, csa_gpu_last_error());
    exit(70);
}

static char *output_name(const char *extra) { /* csamsa.c:41 newOutputFilename */
    int n = (int)strlen(inputfilename), i;
    for (i = n - 1; i > 0; i--)
        if (inputfilename[i] == '.') break;
    if (i == 0) i = n;
    char *r = (char *)calloc((size_t)i + strlen(extra) + 1, 1);
    memcpy(r, inputfilename, (size_t)i);
    strcat(r, extra);
    return r;
}

static int acgt_or_dash(int c) { return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : '-'; }

/* is b a rotation of a (same length), letters outside ACGT all alike? */
static int is_rotation(const char *a, const char *b, int n) {
    char *dbl = (char *)malloc(2 * (size_t)n + 1), *pat = (char *)malloc((size_t)n + 1);
    for (int q = 0; q < 2 * n; q++) dbl[q] = (char)acgt_or_dash(a[q % n]);
    for (int q = 0; q < n; q++) pat[q] = (char)acgt_or_dash(b[q]);
    dbl[2 * n] = pat[n] = 0;
    int hit = strstr(dbl, pat) != NULL;
    free(dbl);
    free(pat);
    return hit;
}

/* nodeslinkedlists.c:128 blockLabel: the blocks of the chain, each spelled as the reference's tree spells it
 * (csa_gpu_batch_block_letters), gaps as dashes (up to 7) or -(n)-; a negative gap steps back over letters
 * already written */
static char *chain_label(int b, const int *depth, const int *interval, const int *next, const char *letters,
                         const long long *off) {
    size_t cap = 256, len = 0;
    char *label = (char *)calloc(cap, 1);
    for (int cur = b; cur != -1; cur = next[cur]) {
        int d = depth[cur];
        while (len + (size_t)d + 32 > cap) {
            label = (char *)realloc(label, cap * 2);
            memset(label + cap, 0, cap);
            cap *= 2;
        }
        memcpy(label + len, letters + off[cur], (size_t)d);
        len += (size_t)d;
        int g = interval[cur];
        if (g < 0) len = ((long long)len + g < 0) ? 0 : len + g;
        else if (g > 7) len += (size_t)sprintf(label + len, "-(%d)-", g);
        else for (int i = 0; i < g; i++) label[len++] = '-';
    }
    label[len] = 0;
    return label;
}

static int chain_is_ring(int b, int nblocks, const int *next) {
    int steps = 0;
    for (int cur = b; cur != -1; cur = next[cur])
        if (++steps > nblocks) return 1;
    return 0;
}

int main(int argc, char **argv) {
    int device = 0;
    const char *dev_env = getenv("CSA_GPU_DEVICE");
    if (dev_env) device = atoi(dev_env);
    printf("%c[%d;%d;%dm[ Multiple Circular Sequence Aligner v1.11 ]%c[0m\n", 0x1B, 1, 31, 47, 0x1B);
    char mode = 0;
    if (argc == 3) {
        mode = argv[1][0];
        if (mode >= 'a' && mode <= 'z') mode = (char)(mode - 32);
    }
    if (mode != 'R') {
        printf("> USAGE:\n\t[   Rotation only    ]\t%s R <multi-fasta-file>\n", argv[0]);
        printf("> (the B200 build covers the rotation path only; use the reference for the other modes)\n> Done!\n");
        return 0;
    }
    inputfilename = argv[2];

    /* ---- LoadSequences ---- */
    printf("> Loading sequences from file <%s> ... ", inputfilename);
    FILE *f = fopen(inputfilename, "r");
    if (!f) die("Sequence file not found");
    fseek(f, 0L, SEEK_END);
    long fsize = ftell(f);
    rewind(f);
    printf("(%ld bytes)\n", fsize);
    char *buf = (char *)malloc((size_t)fsize + 1);
    if (fread(buf, 1, (size_t)fsize, f) != (size_t)fsize) die("Sequence file not found");
    fclose(f);
    int cap = 64, m = 0;
    char **texts = (char **)calloc((size_t)cap, sizeof(char *)), **descs = (char **)calloc((size_t)cap, sizeof(char *));
    int *sizes = (int *)calloc((size_t)cap, sizeof(int));
    long pos = 0;
    while (pos < fsize && buf[pos] != '>') pos++;
    if (pos >= fsize) die("No sequences in file");
    for (;;) {
        while (pos < fsize && buf[pos] != '>') pos++;
        if (pos >= fsize) break;
        long ds = ++pos;
        while (pos < fsize && buf[pos] != '\n' && buf[pos] != '\r') pos++;
        int desclen = (int)(pos - ds);
        char *desc = (char *)calloc((size_t)desclen + 1, 1);
        memcpy(desc, buf + ds, (size_t)desclen);
        if (pos < fsize) pos++;
        printf("# %02d [", m + 1);
        int k = 0;
        while (k < 40 && k < desclen) printf("%c", desc[k++]);
        while (k++ < 40) printf(" ");
        printf("] ");
        long se = pos;
        while (se < fsize && buf[se] != '>') se++;
        char *text = (char *)calloc((size_t)(se - pos) + 1, 1);
        int len = 0, bad = 0;
        for (; pos < se; pos++) {
            int c = (unsigned char)buf[pos];
            if (c == '\n' || c == '\r' || c == 0 || c == '-' || c == ' ') continue;
            if (c >= 'a' && c <= 'z') c -= 32;
            if (c && strchr("ACGTRYSWKMDHBVN", c)) text[len++] = (char)c;
            else { bad = 1; break; }
        }
        if (len == 0 || bad) {
            printf(len == 0 ? "EMPTY\n" : "INVALID_CHARS\n");
            free(desc);
            free(text);
            if (bad) pos++;
            continue;
        }
        printf("OK (%d characters)\n", len);
        if (m == cap) {
            cap *= 2;
            texts = (char **)realloc(texts, sizeof(char *) * (size_t)cap);
            descs = (char **)realloc(descs, sizeof(char *) * (size_t)cap);
            sizes = (int *)realloc(sizes, sizeof(int) * (size_t)cap);
        }
        texts[m] = text; descs[m] = desc; sizes[m] = len; m++;
        if (m == MAXNUMBEROFSEQS && !getenv("CSA_NO_SEQ_LIMIT")) {
            printf("> WARNING: Current version only supports up to %d sequences\n", MAXNUMBEROFSEQS);
            break;
        }
    }
    if (m < 2) die("Not enough valid sequences found");
    printf("> %d sequences successfully loaded\n", m);

    /* ---- buildGeneralizedTree: here only its bookkeeping (rotation duplicates, progress dots) ---- */
    printf("> Building generalized cyclic suffix tree");
    fflush(stdout);
    int *shown_id = (int *)malloc(sizeof(int) * (size_t)m); /* number of the sequence as the user saw it loaded */
    for (int i = 0; i < m; i++) shown_id[i] = i;
    for (int j = 0; j < m; j++) {
        int dup = -1;
        for (int i = 0; i < j && dup < 0; i++)
            if (sizes[i] == sizes[j] && is_rotation(texts[i], texts[j], sizes[j])) dup = i;
        if (dup >= 0) {
            printf("> WARNING: Discarding seq. %d because it is an identical rotation of seq. %d\n", shown_id[j] + 1,
                   shown_id[dup] + 1);
            for (int i = j; i < m; i++) shown_id[i]++;
            free(texts[j]); free(descs[j]);
            for (int i = j + 1; i < m; i++) { texts[i - 1] = texts[i]; descs[i - 1] = descs[i]; sizes[i - 1] = sizes[i]; }
            m--;
            if (m < 2) die("The program needs at least 2 sequences to run");
            j--;
        }
        printf(".");
    }
    printf("\n");
    fflush(stdout);

    /* ---- the hot path: one call ---- */
    /* CSA_GPUS=<n> (n > 1): this one process drives n GPUs on the set -- suffix-array buckets sharded over them,
       exchanged by peer copies (csa_gpu_multi_*); worth it for bacterial-scale sets */
    csa_gpu_ctx *ctx = NULL;
    csa_gpu_multi *multi = NULL;
    int ngpus = getenv("CSA_GPUS") ? atoi(getenv("CSA_GPUS")) : 1;
    int *rotations = (int *)calloc((size_t)m, sizeof(int));
    csa_gpu_set_info info;
    if (ngpus > 1) {
        int set_start[2] = {0, m};
        if (csa_gpu_multi_create(ngpus, NULL, &multi) != CSA_GPU_OK) {
            fail_gpu("not that many CUDA devices");
        }
        ctx = csa_gpu_multi_ctx(multi, 0);
        if (csa_gpu_multi_batch_rotations(multi, 1, set_start, (const char *const *)texts, sizes, INT_MAX, CSA_GPU_FLAG_STATS,
                                          rotations, &info) != CSA_GPU_OK)
            fail_gpu("GPU run failed");
    } else {
        if (csa_gpu_create(device, &ctx) != CSA_GPU_OK) fail_gpu("no CUDA device");
        if (csa_gpu_find_rotations(ctx, m, (const char *const *)texts, sizes, INT_MAX, CSA_GPU_FLAG_STATS, rotations, &info) != CSA_GPU_OK)
            fail_gpu("GPU run failed");
    }
    printf("> Collecting maximum common subsequences... ");
    fflush(stdout);
    if (info.count_collected == 0) die("No common subsequences found");
    printf("%d nodes found\n", info.count_collected);
    printf("> Removing suffixes... ");
    fflush(stdout);
    if (info.status == CSA_SET_UNDEFINED)
        stop(6, "removeSuffixNodes (csamsa.c:80) frees the list item it stands on here; the reference's answer is not defined");
    printf("%d nodes left\n", info.count_suffixfree);
    printf("> Removing repeats... ");
    fflush(stdout);
    if (info.count_unique == 0) die("No unique subsequences found");
    printf("%d nodes left\n", info.count_unique);
    printf("> Connecting block chains... ");
    fflush(stdout);
    if (info.status == CSA_SET_DEGENERATE)
        stop(3, "collectNodeChains walks off a leaf here (csamsa.c:153): a whole rotation of the shortest sequence occurs in all others");
    if (info.status == CSA_SET_NONTERMINATING)
        stop(4, "collectNodeChains (csamsa.c:197) does not terminate on this input: the common blocks form a cycle");
    printf("%d chains found\n", info.count_chains);

    /* ---- saveRotatedSequences ---- */
    char *fn = output_name("-Rotated.fasta");
    FILE *o = fopen(fn, "w");
    if (!o) die("Can't write rotated sequences file");
    for (int i = 0; i < m; i++) {
        int rot = rotations[i];
        fprintf(o, ">%s @ %d\n", descs[i], rot);
        fputs(texts[i] + rot, o);
        fwrite(texts[i], 1, (size_t)rot, o);
        fprintf(o, "\n");
    }
    fclose(o);
    free(fn);

    /* ---- createImageAndShowResults: the text part ---- */
    int nb = info.nblocks;
    int *depth = (int *)calloc((size_t)nb, sizeof(int)), *size = (int *)calloc((size_t)nb, sizeof(int));
    int *total = (int *)calloc((size_t)nb, sizeof(int)), *interval = (int *)calloc((size_t)nb, sizeof(int));
    int *next = (int *)calloc((size_t)nb, sizeof(int)), *positions = (int *)calloc((size_t)nb * (size_t)m, sizeof(int));
    long long *off = (long long *)calloc((size_t)nb + 1, sizeof(long long));
    if (csa_gpu_batch_blocks(ctx, depth, size, total, interval, next, positions) != CSA_GPU_OK) fail_gpu("GPU run failed");
    long long nletters = csa_gpu_batch_block_letters(ctx, NULL, off);
    if (nletters < 0) fail_gpu("GPU run failed");
    char *letters = (char *)calloc((size_t)nletters + 1, 1);
    if (nletters && csa_gpu_batch_block_letters(ctx, letters, NULL) < 0) fail_gpu("GPU run failed");
    fn = output_name("-Blocks.csv");
    o = fopen(fn, "w");
    if (!o) die("Can't write original blocks file");
    free(fn);
    fprintf(o, "Length,Sequence");
    for (int i = 0; i < m; i++) fprintf(o, ",Position_%d", i + 1);
    fprintf(o, "\n");
    const int ntoprint = 20, charstoprint = 100;
    int nchains = 0;
    printf("> Length, sequence and rotations for the first %d longest block chains:\n", ntoprint);
    for (int b = 0; b < nb; b++) {
        if (total[b] == -1) continue;
        if (chain_is_ring(b, nb, next)) {
            fclose(o);
            stop(5, "this block chain closes into a ring: blockLabel (nodeslinkedlists.c:150) never returns in the reference");
        }
        char *s = chain_label(b, depth, interval, next, letters, off);
        if (nchains < ntoprint) {
            printf(":: (%d) ", size[b]);
            if ((int)strlen(s) < charstoprint) printf("%s", s);
            else { fwrite(s, 1, (size_t)charstoprint, stdout); printf("..."); }
            printf("\n");
        }
        fprintf(o, "%d,%s", total[b], s);
        for (int i = 0; i < m; i++) fprintf(o, ",%d", positions[(size_t)b * m + i]);
        fprintf(o, "\n");
        free(s);
        nchains++;
    }
    if (nchains > ntoprint) printf(":: ... (%d total)\n", nchains);
    fclose(o);
    printf("> Done!\n");
    if (multi) csa_gpu_multi_destroy(multi); else csa_gpu_destroy(ctx);
    return 0;
}
