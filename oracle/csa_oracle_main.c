/*
 * csa_oracle_main.c -- `csa_oracle R <multi-fasta>`: the oracle behind the reference's CLI, so
 * that oracle/validate_against_ref.py can diff stdout, <base>-Rotated.fasta and
 * <base>-Blocks.csv byte for byte with the compiled reference.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows csamsa.c:437 LoadSequences, :421 saveRotatedSequences, :361
 * createImageAndShowResults (text outputs only; the .bmp and its -positions/-imagemap
 * side files are drawing code, outside the rotation path).
 */
#include "csa_oracle.h"
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXNUMBEROFSEQS 64 /* csamsa.c:22 */

static char *inputfilename;

static char *new_output_filename(const char *extra) { /* csamsa.c:41 */
    int i, n = (int)strlen(inputfilename);
    for (i = n - 1; i > 0; i--) if (inputfilename[i] == '.') break;
    if (i == 0) i = n;
    char *r = (char *)calloc(i + strlen(extra) + 1, 1);
    strncpy(r, inputfilename, i);
    strcat(r, extra);
    return r;
}

static void exit_message(const char *msg) { /* csamsa.c:57 */
    printf("\n> ERROR: %s\n", msg);
    exit(0);
}

static int is_iupac_upper(int c) { return c && strchr("ACGTRYSWKMDHBVN", c) != NULL; }

int main(int argc, char **argv) {
    int max_seqs = MAXNUMBEROFSEQS;
    if (argc >= 4 && strcmp(argv[3], "--no-seq-limit") == 0) max_seqs = INT_MAX; /* oracle-only */
    if (argc < 3 || (argv[1][0] != 'R' && argv[1][0] != 'r')) {
        fprintf(stderr, "usage: %s R <multi-fasta-file> [--no-seq-limit]\n", argv[0]);
        return 2;
    }
    inputfilename = argv[2];
    printf("%c[%d;%d;%dm[ Multiple Circular Sequence Aligner v1.11 ]%c[0m\n", 0x1B, 1, 31, 47, 0x1B);
    /* ---- csamsa.c:437 LoadSequences ---- */
    printf("> Loading sequences from file <%s> ... ", inputfilename);
    FILE *f = fopen(inputfilename, "r");
    if (!f) exit_message("Sequence file not found");
    fseek(f, 0L, SEEK_END);
    long fsize = ftell(f);
    rewind(f);
    printf("(%ld bytes)\n", fsize);
    char *buf = (char *)malloc(fsize + 1);
    if (fread(buf, 1, fsize, f) != (size_t)fsize) exit_message("Sequence file not found");
    fclose(f);
    int cap = 64, m = 0;
    char **texts = (char **)calloc(cap, sizeof(char *)), **descs = (char **)calloc(cap, sizeof(char *));
    int *sizes = (int *)calloc(cap, sizeof(int));
    long pos = 0;
    while (pos < fsize && buf[pos] != '>') pos++;
    if (pos >= fsize) exit_message("No sequences in file");
    while (1) {
        while (pos < fsize && buf[pos] != '>') pos++;
        if (pos >= fsize) break;
        pos++; /* past '>' */
        long ds = pos;
        while (pos < fsize && buf[pos] != '\n' && buf[pos] != '\r') pos++;
        int desclen = (int)(pos - ds);
        char *desc = (char *)calloc(desclen + 1, 1);
        memcpy(desc, buf + ds, desclen);
        if (pos < fsize) pos++; /* the line terminator was consumed by fgetc */
        printf("# %02d [", m + 1);
        int k = 0;
        while (k < 40 && k < desclen) printf("%c", desc[k++]);
        while (k < 40) { printf(" "); k++; }
        printf("] ");
        long ss = pos, se = pos;
        while (se < fsize && buf[se] != '>') se++;
        char *text = (char *)calloc(se - ss + 1, 1);
        k = 0;
        int bad = 0;
        for (pos = ss; pos < se; pos++) {
            int c = (unsigned char)buf[pos];
            if (c == '\n' || c == '\r' || c == '\0' || c == '-' || c == ' ') continue;
            if (c >= 'a' && c <= 'z') c -= 32;
            if (is_iupac_upper(c)) text[k++] = (char)c;
            else { bad = 1; break; }
        }
        if (k == 0) { printf("EMPTY\n"); free(desc); free(text); if (bad) pos++; continue; }
        if (bad) { printf("INVALID_CHARS\n"); free(desc); free(text); pos++; continue; }
        printf("OK (%d characters)\n", k);
        if (m == cap) {
            cap *= 2;
            texts = (char **)realloc(texts, sizeof(char *) * cap);
            descs = (char **)realloc(descs, sizeof(char *) * cap);
            sizes = (int *)realloc(sizes, sizeof(int) * cap);
        }
        texts[m] = text; descs[m] = desc; sizes[m] = k; m++;
        if (m == max_seqs) {
            printf("> WARNING: Current version only supports up to %d sequences\n", MAXNUMBEROFSEQS);
            break;
        }
    }
    if (m < 2) exit_message("Not enough valid sequences found");
    printf("> %d sequences successfully loaded\n", m);
    /* ---- gencycsuffixtrees.c:518-524: a sequence that is a rotation of an earlier one is dropped */
    printf("> Building generalized cyclic suffix tree");
    {
        int *oldid = (int *)malloc(sizeof(int) * m);
        for (int i = 0; i < m; i++) oldid[i] = i;
        for (int j = 0; j < m; j++) {
            int dup = -1;
            for (int i = 0; i < j && dup < 0; i++) {
                if (sizes[i] != sizes[j]) continue;
                int n = sizes[j];
                char *dbl = (char *)malloc(2 * n + 1);
                for (int q = 0; q < 2 * n; q++) {
                    char c = texts[i][q % n];
                    dbl[q] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : '-';
                }
                dbl[2 * n] = 0;
                char *pat = (char *)malloc(n + 1);
                for (int q = 0; q < n; q++) {
                    char c = texts[j][q];
                    pat[q] = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : '-';
                }
                pat[n] = 0;
                if (strstr(dbl, pat)) dup = i;
                free(dbl); free(pat);
            }
            if (dup >= 0) {
                printf("> WARNING: Discarding seq. %d because it is an identical rotation of seq. %d\n",
                       oldid[j] + 1, oldid[dup] + 1);
                for (int i = j; i < m; i++) oldid[i]++;
                free(texts[j]); free(descs[j]);
                for (int i = j + 1; i < m; i++) {
                    texts[i - 1] = texts[i]; descs[i - 1] = descs[i]; sizes[i - 1] = sizes[i];
                }
                m--;
                if (m < 2) exit_message("The program needs at least 2 sequences to run");
                j--;
            }
            printf("."); /* gencycsuffixtrees.c:540 runs after the break too */
        }
        free(oldid);
    }
    printf("\n");
    csa_oracle_result r;
    csa_oracle_run(m, (const char *const *)texts, sizes, INT_MAX, &r);
    printf("> Collecting maximum common subsequences... ");
    if (r.count_collected == 0) exit_message("No common subsequences found");
    printf("%d nodes found\n", r.count_collected);
    printf("> Removing suffixes... ");
    fflush(stdout);
    if (r.status == CSA_ORACLE_UNDEFINED) {
        printf("\n> ORACLE: removeSuffixNodes (csamsa.c:80) frees the list item it stands on; what the reference does next is not defined\n");
        return 6;
    }
    printf("%d nodes left\n", r.count_suffixfree);
    printf("> Removing repeats... ");
    if (r.count_unique == 0) exit_message("No unique subsequences found");
    printf("%d nodes left\n", r.count_unique);
    printf("> Connecting block chains... ");
    fflush(stdout);
    /* exit codes 3, 4, 5: the oracle's statement that the reference does not get past this point
     * (oracle/validate_against_ref.py checks that it indeed dies or hangs, and only then) */
    if (r.status == CSA_ORACLE_DEGENERATE) {
        printf("\n> ORACLE: the reference walks off a leaf here (csamsa.c:153; a whole rotation of the shortest sequence occurs in all others)\n");
        return 3;
    }
    if (r.status == CSA_ORACLE_HANG) {
        printf("\n> ORACLE: the reference does not terminate on this input (block cycle, csamsa.c:197)\n");
        return 4;
    }
    printf("%d chains found\n", r.count_chains);
    /* ---- csamsa.c:421 saveRotatedSequences ---- */
    char *fn = new_output_filename("-Rotated.fasta");
    FILE *o = fopen(fn, "w");
    if (!o) exit_message("Can't write rotated sequences file");
    for (int i = 0; i < m; i++) {
        int rot = r.rotations[i];
        fprintf(o, ">%s @ %d\n", descs[i], rot);
        fputs(texts[i] + rot, o);
        fwrite(texts[i], 1, (size_t)rot, o);
        fprintf(o, "\n");
    }
    fclose(o);
    free(fn);
    /* ---- csamsa.c:361 createImageAndShowResults (Blocks.csv + console list) ---- */
    fn = new_output_filename("-Blocks.csv");
    o = fopen(fn, "w");
    if (!o) exit_message("Can't write original blocks file");
    free(fn);
    fprintf(o, "Length,Sequence");
    for (int i = 0; i < m; i++) fprintf(o, ",Position_%d", i + 1);
    fprintf(o, "\n");
    int ntoprint = 20, charstoprint = 100, nchains = 0;
    printf("> Length, sequence and rotations for the first %d longest block chains:\n", ntoprint);
    for (int b = 0; b < r.nblocks; b++) {
        if (r.totalsize[b] == -1) continue;
        if (csa_oracle_chain_is_ring(&r, b)) {
            fflush(stdout);
            fclose(o);
            printf("> ORACLE: this chain closes into a ring; blockLabel (nodeslinkedlists.c:150) never returns\n");
            return 5;
        }
        char *s = csa_oracle_block_label(&r, b);
        if (nchains < ntoprint) {
            printf(":: (%d) ", r.size[b]);
            if ((int)strlen(s) < charstoprint) printf("%s", s);
            else { for (int i = 0; i < charstoprint; i++) printf("%c", s[i]); printf("..."); }
            printf("\n");
        }
        fprintf(o, "%d,%s", r.totalsize[b], s);
        for (int i = 0; i < m; i++) fprintf(o, ",%d", r.positions[(size_t)b * m + i]);
        fprintf(o, "\n");
        free(s);
        nchains++;
    }
    if (nchains > ntoprint) printf(":: ... (%d total)\n", nchains);
    fclose(o);
    printf("> Done!\n");
    return 0;
}
