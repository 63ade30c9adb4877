/*
 * csa_oracle.h -- CPU restatement of fjdf/CSA's rotation-finding path (`./CSA R`).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, link or execute it, and only as the checker.
 *
 * Parity status: PINNED.  oracle/validate_against_ref.py runs this restatement and the
 * unmodified reference (compiled from /root/reference/source by oracle/Makefile into
 * oracle/_ref/CSA_ref) on Manual/Primates.txt, Manual/Mammals.txt and thousands of
 * seeded synthetic sets and requires byte-identical -Rotated.fasta, -Blocks.csv and
 * stdout counts.  The vectors that travel to the GPU box live in tests/golden/.
 *
 * The reference answers the question with a generalized cyclic suffix tree
 * (gencycsuffixtrees.c:418 buildGeneralizedTree) and linked lists of tree nodes
 * (csamsa.c:324 analyzeTree).  This file restates the same mathematics on a
 * generalized cyclic suffix array + LCP array, which is what the CUDA path builds too,
 * and keeps every tie-break of the reference (see csa_oracle.c for file:line notes).
 */
#ifndef CSA_ORACLE_H
#define CSA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
    CSA_ORACLE_OK = 0,
    CSA_ORACLE_NO_COMMON = 1,   /* csamsa.c:330 "No common subsequences found"      */
    CSA_ORACLE_NO_UNIQUE = 2,   /* csamsa.c:346 "No unique subsequences found"       */
    CSA_ORACLE_DEGENERATE = 3,  /* collectNodeChains' walk (csamsa.c:147-183) reaches a LEAF that holds
                                   every sequence (a whole rotation of the shortest one inside all
                                   others): the reference dereferences NULL or loops there          */
    CSA_ORACLE_HANG = 4,        /* reference would not terminate (zero-gap block cycle) */
    CSA_ORACLE_UNDEFINED = 5    /* removeSuffixNodes (csamsa.c:80) frees the list item it stands on and
                                   reads it afterwards (a sequence w^c whose rotations, followed round,
                                   come back to the one the walk started from)                        */
};

typedef struct csa_oracle_result {
    int status;
    int m;                 /* number of sequences                                         */
    int count_collected;   /* "%d nodes found"  csamsa.c:332                               */
    int count_suffixfree;  /* "%d nodes left" after removeSuffixNodes  csamsa.c:338        */
    int count_unique;      /* "%d nodes left" after removeNonUniqueNodes csamsa.c:348      */
    int count_chains;      /* "%d chains found" csamsa.c:354                               */
    int nblocks;           /* blocks in the final (sorted) list == count_unique            */
    /* per block, in the order of the reference's final blockslist (after sortList)       */
    int *depth;            /* item->depth                                                  */
    int *size;             /* linkedblock.size                                             */
    int *totalsize;        /* linkedblock.totalsize (-1: not a chain head)                 */
    int *interval;         /* linkedblock.interval                                         */
    int *next;             /* index (in this list) of nextblock, -1 if none                */
    int *positions;        /* nblocks x m, positions[b*m+k]                                */
    int *rotations;        /* m; NULL unless status==OK                                    */
    char **letters;        /* per block: its depth letters as nodeslinkedlists.c:154-165 spells them */
} csa_oracle_result;

/* texts: upper-case IUPAC letters as produced by the reference loader (csamsa.c:517);
 * anything that is not A/C/G/T compares equal to any other such letter
 * (gencycsuffixtrees.c:283).  max_interval: csamsa.c:27 (INT_MAX on the R path). */
int csa_oracle_run(int m, const char *const *texts, const int *textsizes, int max_interval,
                   csa_oracle_result *out);
int csa_oracle_run_sa(int m, const char *const *texts, const int *textsizes, int max_interval,
                      csa_oracle_result *out, int *sa_out, int *lcp_out);
void csa_oracle_free(csa_oracle_result *r);

/* chain label exactly as nodeslinkedlists.c:144 blockLabel prints it (malloc'd). */
char *csa_oracle_block_label(const csa_oracle_result *r, int b);
/* 1 when the chain starting at block b closes into a ring (blockLabel never returns: the reference dies) */
int csa_oracle_chain_is_ring(const csa_oracle_result *r, int b);

/* flat helpers for ctypes users (tests/, bench.py): the generalized cyclic suffix array
 * and LCP of one set, global index = offset[k]+p.  Returns 0 on success. */
int csa_oracle_gsa(int m, const char *const *texts, const int *textsizes, int *sa_out,
                   int *lcp_out);

#ifdef __cplusplus
}
#endif
#endif
