/*
 * csa_gpu.h -- C ABI of the B200 rotation finder (libcsa_gpu.so).
 *
 * Drop-in boundary for fjdf/CSA's `./CSA R <multi-fasta>` hot path.  In the reference that
 * path is two calls on globals (csamsa.c:599 and :610):
 *
 *     tree = buildGeneralizedTree();   // gencycsuffixtrees.c:418  generalized cyclic suffix tree
 *     analyzeTree();                   // csamsa.c:324  collectNodes / removeSuffixNodes /
 *                                      //   removeNonUniqueNodes / collectNodeChains / getRotations
 *
 * whose inputs are `numberofseqs`, `texts`, `textsizes` (csamsa.h:8-11) and whose outputs are
 * `rotations` (csamsa.h:12) and the sorted `blockslist` (csamsa.c:30) that
 * saveRotatedSequences (csamsa.c:421) and createImageAndShowResults (csamsa.c:361) print.
 * csa_gpu_find_rotations() below replaces exactly those two calls; the batch entry points
 * run many independent sequence sets in one pass and are what bench.py and the sharded
 * (one process per GPU) driver use.  INTEGRATION.md shows the patch to csamsa.c.
 *
 * Plain pointers and sizes only.  Every function returns CSA_GPU_OK (0) or a negative
 * CSA_GPU_E* code; csa_gpu_last_error() gives the text.  There is no CPU fallback: without
 * a CUDA device csa_gpu_create fails with CSA_GPU_ENODEV.
 */
#ifndef CSA_GPU_H
#define CSA_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define CSA_GPU_ABI_VERSION 1

/* return codes */
#define CSA_GPU_OK 0
#define CSA_GPU_ENODEV (-1)   /* no usable CUDA device                      */
#define CSA_GPU_EINVAL (-2)   /* bad argument                               */
#define CSA_GPU_ENOMEM (-3)   /* device or host allocation failed           */
#define CSA_GPU_ECUDA (-4)    /* a CUDA call or kernel failed               */
#define CSA_GPU_ESTATE (-5)   /* call order (run before upload, ...)        */

/* per-set status (csa_gpu_set_info.status).  The first three mirror the reference's exits. */
#define CSA_SET_OK 0
#define CSA_SET_NO_COMMON 1      /* csamsa.c:330 "No common subsequences found"              */
#define CSA_SET_NO_UNIQUE 2      /* csamsa.c:346 "No unique subsequences found"              */
#define CSA_SET_DEGENERATE 3     /* collectNodeChains' walk (csamsa.c:147-183) reaches a LEAF that holds every
                                    sequence (a whole rotation of the shortest one inside all others): the
                                    reference dereferences NULL (csamsa.c:153) or loops there.  A set with such
                                    a leaf that the walk does not reach is answered like any other         */
#define CSA_SET_NONTERMINATING 4 /* the reference loops forever in collectNodeChains
                                    (csamsa.c:197) on a block cycle whose gaps sum to <= 0   */

#define CSA_SET_UNDEFINED 5      /* removeSuffixNodes (csamsa.c:80) frees the list item it stands on and reads
                                    it afterwards (a sequence w^c whose rotations, followed round, lead back
                                    to the one the walk started from): the reference's answer is not defined  */

/* flags for csa_gpu_batch_run */
#define CSA_GPU_FLAG_STATS 1u /* also count "nodes found"/"nodes left" (csamsa.c:332,338)    */

typedef struct csa_gpu_ctx csa_gpu_ctx;

typedef struct csa_gpu_set_info {
    int status;           /* CSA_SET_*                                                       */
    int nseqs;            /* sequences in the set                                            */
    int count_collected;  /* csamsa.c:332 (-1 unless CSA_GPU_FLAG_STATS)                     */
    int count_suffixfree; /* csamsa.c:338 (-1 unless CSA_GPU_FLAG_STATS)                     */
    int count_unique;     /* csamsa.c:348 == nblocks                                         */
    int count_chains;     /* csamsa.c:354                                                    */
    int nblocks;          /* blocks in the final sorted blockslist                           */
    int chain_is_cyclic;  /* 1: the head chain loops back on itself; the reference then
                             overruns blockLabel's buffer (nodeslinkedlists.c:161) after it
                             has written -Rotated.fasta                                      */
    long long block_offset; /* first block of this set in the batch-wide block arrays        */
} csa_gpu_set_info;

/* ---- lifetime ------------------------------------------------------------------------ */
int csa_gpu_create(int device, csa_gpu_ctx **ctx);
void csa_gpu_destroy(csa_gpu_ctx *ctx);
const char *csa_gpu_last_error(void);
int csa_gpu_abi_version(void);
/* run on the caller's CUDA stream (a cudaStream_t, e.g. the host framework's current stream)
 * instead of the context's own; NULL goes back to the context's stream */
int csa_gpu_set_stream(csa_gpu_ctx *ctx, void *cuda_stream);

/* ---- batch of independent sequence sets ---------------------------------------------- */
/* Sequences are given flat: sequence i has texts[i][0..textsizes[i]) -- upper-case IUPAC
 * letters exactly as the reference loader stores them (csamsa.c:517-523); anything that is
 * not A/C/G/T is one and the same fifth letter (gencycsuffixtrees.c:283).  Set s owns
 * sequences [set_start[s], set_start[s+1]).  Sequences that are rotations of an earlier one
 * must already have been dropped (the host layer does it, as gencycsuffixtrees.c:518). */
int csa_gpu_batch_upload(csa_gpu_ctx *ctx, int nsets, const int *set_start,
                         const char *const *texts, const int *textsizes);
/* same, from one contiguous buffer: sequence i is text[text_start[i] .. text_start[i+1]) */
int csa_gpu_batch_upload_flat(csa_gpu_ctx *ctx, int nsets, const int *set_start,
                              const char *text, const long long *text_start);
/* page-lock / release a host buffer (cudaHostRegister): csa_gpu_batch_upload_flat copies straight
 * from a page-locked `text`, without the staging copy it otherwise makes */
int csa_gpu_pin_host(const void *p, unsigned long long bytes);
int csa_gpu_unpin_host(const void *p);
/* all kernels; inputs and outputs stay in HBM.  max_interval: csamsa.c:27 (INT_MAX on R). */
int csa_gpu_batch_run(csa_gpu_ctx *ctx, int max_interval, unsigned flags);
/* rotations: one int per sequence of the batch (csamsa.h:12); info: one per set. */
int csa_gpu_batch_download(csa_gpu_ctx *ctx, int *rotations, csa_gpu_set_info *info);
/* total number of blocks of the last run, and of block positions (sum over blocks of the number
 * of sequences of the block's set), for sizing the arrays below */
long long csa_gpu_batch_num_blocks(csa_gpu_ctx *ctx);
long long csa_gpu_batch_num_positions(csa_gpu_ctx *ctx);
/* the sorted blockslist of every set (set s: blocks [block_offset, block_offset+nblocks)):
 * depth/size/totalsize/interval as in nodeslinkedlists.h:4-13, next = index of nextblock
 * within the set's list or -1, positions: per block, one int per sequence of its set, stored
 * at positions[position_offset(block)] where position_offset is the running sum of nseqs.
 * Any pointer may be NULL. */
int csa_gpu_batch_blocks(csa_gpu_ctx *ctx, int *depth, int *size, int *totalsize, int *interval,
                         int *next, int *positions);
/* the letters of every block as blockLabel (nodeslinkedlists.c:128-165) spells them: it reads each tree edge's
 * label from the text that CREATED the edge (labelfrom/startpos), so a letter outside ACGT -- all of which
 * compare equal -- is printed as the FIRST occurrence in sequence 0 of the block's prefix up to that letter
 * has it, not as the block's own place has it.  Block b (batch-wide index, final list order) owns
 * letters[offsets[b] .. offsets[b+1]); offsets has num_blocks+1 entries.  Returns the total number of letters
 * (call with letters == NULL to size the buffer), or a negative CSA_GPU_E* code. */
long long csa_gpu_batch_block_letters(csa_gpu_ctx *ctx, char *letters, long long *offsets);
/* upload + run + download in one call with host buffers (what bench.py times as e2e) */
int csa_gpu_batch_rotations(csa_gpu_ctx *ctx, int nsets, const int *set_start,
                            const char *const *texts, const int *textsizes, int max_interval,
                            unsigned flags, int *rotations, csa_gpu_set_info *info);

/* ---- the reference's call, one set ---------------------------------------------------- */
/* Replaces buildGeneralizedTree()+analyzeTree(): fills rotations[numberofseqs] and *info. */
int csa_gpu_find_rotations(csa_gpu_ctx *ctx, int numberofseqs, const char *const *texts,
                           const int *textsizes, int max_interval, unsigned flags, int *rotations,
                           csa_gpu_set_info *info);

/* ---- introspection used by tests and bench.py ----------------------------------------- */
/* generalized cyclic suffix array + LCP of the last run, global suffix index = offset of the
 * sequence in the batch + position; n = total number of bases.  Either pointer may be NULL. */
long long csa_gpu_batch_num_suffixes(csa_gpu_ctx *ctx);
int csa_gpu_batch_suffix_array(csa_gpu_ctx *ctx, unsigned *sa, int *lcp);
/* device time of the stages of the last run, in milliseconds (CUDA events on the run's
 * stream): [0] suffix sort, [1] LCP, [2] block discovery, [3] block order, [4] chaining,
 * [5] whole run; launches = kernels launched by the last run. */
int csa_gpu_batch_timings(csa_gpu_ctx *ctx, float ms[6], long long *launches);

/* ---- one batch, the suffix-array stage sharded over the ranks of a job ----------------------------------
 * For a single set too large to be quick on one GPU (BASELINE configs[4]: bacterial chromosomes).  The
 * reference builds ONE generalized tree (gencycsuffixtrees.c:418 buildGeneralizedTree); its counterpart here,
 * the suffix array, splits by first letters into buckets that are ordered independently of one another:
 *   every rank:  csa_gpu_batch_upload* of the same batch, then
 *   csa_gpu_shard_begin(ctx, rank, nranks)   first sort (the same on every rank), bucket borders at group
 *                                             borders, groups + LCPs of bucket `rank` by the word sort
 *   csa_gpu_shard_view(ctx, &v)              device pointers and the borders: the caller copies
 *                                             sa/head/lcp[bounds[r] .. bounds[r+1]) from rank r to all ranks
 *                                             (NCCL broadcast / all-gather), concatenates the ranks' `left` lists
 *                                             into `left` and merges nleft (sum), left_suffixes (sum), min_depth
 *                                             (min), max_group (max)
 *   csa_gpu_shard_finish(ctx, ...merged...)  the rest of the path on every rank (results as csa_gpu_batch_run)
 * No collective runs inside the library; csa_b200/shard.py drives the exchange with torch.distributed. */
typedef struct csa_gpu_shard_info {
    void *sa, *head, *lcp;        /* device, unsigned[n] */
    void *left;                   /* device, unsigned long long[]: groups the bucket sort left (start << 32 | size) */
    unsigned long long n;
    const unsigned *bounds;       /* host, nranks + 1 entries, valid until the next call on ctx */
    unsigned nleft, left_suffixes, min_depth, max_group;
    unsigned own_sort;            /* 1: this rank sorted only its own bucket (one set); 0: the first sort ran on every rank */
} csa_gpu_shard_info;
int csa_gpu_shard_begin(csa_gpu_ctx *ctx, int rank, int nranks);
/* Does sharding the uploaded batch over nranks ranks by buckets pay?  1 yes, 0 no: ONE set of whole genomes is ordered
 * on one GPU by carrying a column's order over to the next along the text (pipeline.cuh "carried word sort"), and the
 * walks cross the buckets at every step -- a bucket's rank must order most of its groups letter by letter again.  That
 * beats the one-GPU run only from CSA_GPU_SHARD_MIN_RANKS ranks on; below, every rank had better run the whole set
 * (csa_gpu_batch_run; "replicas").  Batches of several sets always shard (no walk leaves its set's stretch). */
#define CSA_GPU_SHARD_MIN_RANKS 6
int csa_gpu_shard_advice(csa_gpu_ctx *ctx, int nranks);
int csa_gpu_shard_view(csa_gpu_ctx *ctx, csa_gpu_shard_info *out);
int csa_gpu_shard_finish(csa_gpu_ctx *ctx, int max_interval, unsigned flags, unsigned nleft, unsigned left_suffixes,
                         unsigned min_depth, unsigned max_group);
/* ---- ... and the block stages too (one set, every bucket finished by its rank: nleft == 0 everywhere) ----------------------
 * Instead of copying every bucket of the suffix array and the LCP array to every rank (2 x 4 N bytes) and finding the
 * blocks on every rank, each rank finds the blocks that START in its own range and picks out its rotations of sequence 0;
 * what travels is a record per block (place, depth, the m suffixes of its window) and sequence 0's three arrays:
 *   caller:  copy sa and lcp [bounds[r+1] .. bounds[r+1]+m) from rank r+1 and sa[bounds[r]-1] from rank r-1 into rank r's
 *            arrays (csa_gpu_shard_view: the same global places); needs every bucket to hold at least m places
 *   csa_gpu_shard_blocks_begin(ctx, flags, &b)   this range's block records and rotations of sequence 0
 *   caller:  all-gather the counts; if some rank reports b.rare (a full-length match: csa_b200/csrc/rare.cuh redoes such
 *            sets one thread per set) fall back on the full exchange + csa_gpu_shard_finish; else
 *            csa_gpu_shard_blocks_buffers(ctx, total blocks, total n0, ...) and gather the ranks' records / arrays into
 *            them in rank order; lcp0 of the first rotation of a rank's range = min(its head_min, the tail_min of the
 *            ranks before it back to the last one that holds a rotation of sequence 0) (0 for the very first)
 *   csa_gpu_shard_blocks_finish(ctx, ...)        block order, chaining, rotations on every rank (results as csa_gpu_batch_run;
 *            csa_gpu_batch_suffix_array is then valid for the rank's own range and the blocks' windows only) */
typedef struct csa_gpu_shard_blocks {
    void *blkrec;                 /* device, unsigned[nblk][2 + m] */
    unsigned nblk, m;
    void *sa0, *saidx0, *lcp0;    /* device, unsigned[n0]: position in sequence 0, SA place, LCP with the rotation before */
    unsigned n0;
    unsigned head_min, tail_min;  /* smallest lcp from the range's start to its first rotation of sequence 0 / behind its last
                                     one to the range's end (0xFFFFFFFF: none); n0 == 0: head_min = the whole range */
    unsigned rare;
} csa_gpu_shard_blocks;
int csa_gpu_shard_blocks_begin(csa_gpu_ctx *ctx, unsigned flags, csa_gpu_shard_blocks *out);
int csa_gpu_shard_blocks_buffers(csa_gpu_ctx *ctx, unsigned total_blocks, unsigned total_n0, void **blkrec, void **sa0,
                                 void **saidx0, void **lcp0);
int csa_gpu_shard_blocks_finish(csa_gpu_ctx *ctx, int max_interval, unsigned flags, unsigned total_blocks, unsigned total_n0);

/* The same from ONE process that drives several GPUs (what the C host does: no NCCL, no Python).  One host thread
 * per GPU for the compute phases, the bucket exchange by peer copies.  Results are read from context 0
 * (csa_gpu_multi_ctx(m, 0)) with the csa_gpu_batch_* calls above. */
typedef struct csa_gpu_multi csa_gpu_multi;
int csa_gpu_multi_create(int ngpus, const int *devices /* NULL: 0 .. ngpus-1 */, csa_gpu_multi **out);
void csa_gpu_multi_destroy(csa_gpu_multi *m);
int csa_gpu_multi_size(csa_gpu_multi *m);
csa_gpu_ctx *csa_gpu_multi_ctx(csa_gpu_multi *m, int i);
int csa_gpu_multi_batch_rotations(csa_gpu_multi *m, int nsets, const int *set_start, const char *const *texts,
                                  const int *textsizes, int max_interval, unsigned flags, int *rotations,
                                  csa_gpu_set_info *info);

/* tests: which of the equivalent ways the suffix-array stage takes (they must all give the same suffix array and
 * LCP array).  mode: -1 leaves the setting; 0 free choice (word sort when the groups of the first sort are small,
 * rank doubling otherwise); 1 device-wide radix rounds only; 2 tile rounds with doubling only (+ text-order LCP);
 * 3 tile rounds, quadrupling allowed; 4 word sort stopped after two words, the rest by doubling rounds; 5 free choice
 * among the doubling rounds (no word sort); 6 word sort whatever the groups look like; 7 free choice, long block
 * lists chained by the literal one-thread walk; 8 free choice, sharded runs of one set sort the whole set on every
 * rank; 9 free choice, blocks always found through the cover array (as when the counts are asked for); 10 word sort
 * with the order of a column carried over to the next (what sets of whole genomes take) whatever the sets look like;
 * 11 free choice, but the word sort never carried; 12 as 10 with the group table laid out as for batches of 2^27
 * suffixes and more.  rounds[0..1] = rounds of the last run on the tile/list/word-sort paths and on the device-wide path */
int csa_gpu_debug_rounds(csa_gpu_ctx *ctx, int mode, int rounds[2]);
/* per-kernel profile: with it enabled every launch of the next runs is bracketed by CUDA events on
 * the run's stream; after a run row i gives the kernel's name, its launches, their summed device
 * time (ms) and the ALGORITHMIC bytes they moved (DESIGN.md lists the per-item figures).  Off by
 * default; bench.py enables it for a separate pass, never inside the timed steps. */
int csa_gpu_profile_enable(csa_gpu_ctx *ctx, int on);
int csa_gpu_profile_count(csa_gpu_ctx *ctx);
int csa_gpu_profile_get(csa_gpu_ctx *ctx, int i, char *name, int name_cap, long long *launches, double *ms,
                        double *bytes);

#ifdef __cplusplus
}
#endif
#endif
