// csa_gpu.cu -- C ABI of the B200 rotation finder (include/csa_gpu.h) and the host-side
// sequencing of the kernels in pipeline.cuh.  Built by nvcc for sm_100a into libcsa_gpu.so.
// (With -DCSA_EMU and g++ the same file gives tests/emu/libcsa_emu.so, a CPU single-stepper of
// the kernel bodies used by the CPU-only tests; it is never part of the product.)
#include "pipeline.cuh"
#include "rare.cuh"
#include "../../include/csa_gpu.h"
#include <algorithm>
#include <vector>
#include <string>
#include <new>
#include <thread>

thread_local char g_csa_err[512] = "";

#define TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

static inline int bits_for(u64 x) { // bits needed to hold values 0..x
    int b = 1;
    while (b < 64 && (x >> b)) b++;
    return b;
}

struct StageTimer {
#ifndef CSA_EMU
    cudaEvent_t ev[8];
    bool ok = false;
#endif
    float ms[6] = {0, 0, 0, 0, 0, 0};
};

struct csa_gpu_ctx {
    int device = 0;
    Exec ex{};
    csaStream_t own_stream{};
    PrimScratch ps;
    StageTimer tm;
    // ---- batch description (host) ----
    bool uploaded = false, ran = false;
    int nsets = 0;
    u32 M = 0, N = 0, N0 = 0, nmax = 0, n0max = 0, mmax = 0;
    u64 TW = 0; // words of the doubled text
    std::vector<u32> h_seq_off, h_seq_set, h_set_seq0, h_set_base0, h_set_nmin, h_seq_nmin, h_z0;
    std::vector<u64> h_dbl_off;
    // ---- results (host mirrors of the small arrays) ----
    std::vector<u32> h_set_nblocks, h_set_blk0, h_set_pos0, h_set_flags, h_set_nchains, h_set_cyclic;
    u32 B = 0, E = 0;
    long long launches = 0;
    // ---- device ----
    DevMem raw, code, seqof, p2, pm, seq_off, seq_set, set_seq0, set_base0, set_nmin, seq_nmin, dbl_off, z0;
    DevMem rs_start, rs_count, rs_cbase, rs_stride;
    u32 rs_nblocks = 0;
    DevMem keysA, keysB, valsA, valsB, sa, t0, t1, t2, t3, t4, t5, counter, tiles;
    u32 batch_nmin = 0, max_set_bases = 0;
    int lcp_state = 0;          // after the suffix array stage: 0 nothing known, 1 every LCP known, 2 all but the LCP_UNKNOWN places
    int ws_runs = 0, ws_force = 0; double ws_pairs = 0, ws_sharing = 0;
    u32 sa_any_other = 1, sa_ngroups = 0;
    int shard_rank = 0, shard_nranks = 1, shard_phase = 0, shard_full_sort = 0; bool shard_own_sort = false;
    DevMem bk_hist;
    DevMem shard_bounds; std::vector<u32> h_shard_bounds;
    DevMem chb_sets, chb_evbase, chb_events, chb_work, chb_redo;
    bool use_cover = true; int force_cover = 0; // blocks through the cover array R[] (counts asked for, sets of > 256 sequences) or straight from the LCP array
    int no_chain_big = 0, chain_redone = 0; u32 ws_depth_cap = WS_DEPTH_CAP; u32 ws_left[6] = {0, 0, 0, 0, 0, 0};
    double lcp_mean_sample = 0;
    int force_kasai = 0;
    int rounds_list = 0, round_mode = 0;
    int carry_mode = 0; bool ws_carried = false, carry_pick = false, cy_nopack = false, carry_big = false; // carried word sort: 0 for sets of whole genomes, 1 always, 2 never
    int rounds_tiled = 0, rounds_global = 0, rounds_quad = 0, force_global_rounds = 0, no_quad_rounds = 0;
    DevMem pyr, pyr2, sa0, saidx0, leaf_set, lcp0;
    Seq0Q q0{};                 // sequence 0 of every set (stage_seq0)
    DevMem set_nblocks, set_blk0, set_pos0, set_flags, set_nchains, set_cyclic, firstmax, set_collected, set_suffixfree;
    DevMem set_neff, seq_per, rare_collected, rare_suffixfree; // rare.cuh
    bool have_stats = false;
    std::vector<u32> h_set_collected, h_set_suffixfree;
    DevMem blk_lb, blk_depth, blk_set, order, o_depth, o_set, o_pos, elem_blk, seghead, succ_lo, succ_hi;
    DevMem next, gap, size, total, interval, inv, f_depth, f_size, f_total, f_interval, f_next, f_pos, rotations;
    DevMem blk_leaf, blk_tab, f_leaf, f_set, let_off, let_out; // block order; csa_gpu_batch_block_letters
    DevMem sh_rec, sh_sa0, sh_saidx0, sh_lcp0, sh_allrec;      // csa_gpu_shard_blocks_*: this range's records, everybody's records
    bool shard_sa_swapped = false;
    void *pinned = nullptr, *meta_pinned = nullptr;
    size_t pinned_bytes = 0, meta_bytes = 0;
    // per-kernel profile of the last run (csa_gpu_profile_*)
    Profiler prof;
    struct ProfSum { std::string name; long long launches; double ms, bytes; };
    std::vector<ProfSum> prof_sum;
};

template <class T> static inline T *P(DevMem &m) { return (T *)m.p; }

static BatchView view_of(csa_gpu_ctx *c) {
    BatchView v;
    v.nsets = c->nsets; v.M = c->M; v.N = c->N;
    v.seq_off = P<u32>(c->seq_off); v.seq_set = P<u32>(c->seq_set);
    v.set_seq0 = P<u32>(c->set_seq0); v.set_base0 = P<u32>(c->set_base0);
    v.set_nmin = P<u32>(c->set_nmin); v.seq_nmin = P<u32>(c->seq_nmin); v.dbl_off = P<u64>(c->dbl_off);
    v.seqof = P<u32>(c->seqof); v.code = P<unsigned char>(c->code);
    v.p2 = P<u64>(c->p2); v.pm = P<u32>(c->pm);
    return v;
}

// ---- lifetime ---------------------------------------------------------------------------------
extern "C" int csa_gpu_abi_version(void) { return CSA_GPU_ABI_VERSION; }
extern "C" const char *csa_gpu_last_error(void) { return g_csa_err; }

extern "C" int csa_gpu_create(int device, csa_gpu_ctx **out) {
    if (!out) CSA_FAIL(CSA_GPU_EINVAL, "csa_gpu_create: null ctx pointer");
    *out = nullptr;
#ifndef CSA_EMU
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        CSA_FAIL(CSA_GPU_ENODEV, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) CSA_FAIL(CSA_GPU_ENODEV, "device %d not present (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
#endif
    csa_gpu_ctx *c = new (std::nothrow) csa_gpu_ctx();
    if (!c) CSA_FAIL(CSA_GPU_ENOMEM, "out of host memory");
    c->device = device;
#ifndef CSA_EMU
    {   // (a failure half way must not leak the context, its stream or the events made so far)
        cudaError_t err = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
        int made = 0;
        if (err == cudaSuccess) {
            c->ex.stream = c->own_stream;
            for (; made < 8 && err == cudaSuccess; made++) err = cudaEventCreate(&c->tm.ev[made]);
            if (err != cudaSuccess) made--;
        }
        if (err != cudaSuccess) {
            for (int i = 0; i < made; i++) cudaEventDestroy(c->tm.ev[i]);
            if (c->ex.stream) cudaStreamDestroy(c->own_stream);
            delete c;
            CSA_FAIL(err == cudaErrorMemoryAllocation ? CSA_GPU_ENOMEM : CSA_GPU_ECUDA, "csa_gpu_create: %s", cudaGetErrorString(err));
        }
        c->tm.ok = true;
    }
#endif
    *out = c;
    return CSA_GPU_OK;
}

// run on the caller's stream (e.g. torch's current stream) instead of the context's own
extern "C" int csa_gpu_set_stream(csa_gpu_ctx *c, void *stream) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
#ifndef CSA_EMU
    CUDA_TRY(cudaStreamSynchronize(c->ex.stream));
    c->ex.stream = stream ? (cudaStream_t)stream : c->own_stream;
#else
    (void)stream;
#endif
    return CSA_GPU_OK;
}

extern "C" void csa_gpu_destroy(csa_gpu_ctx *c) {
    if (!c) return;
#ifndef CSA_EMU
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->ex.stream);
#endif
    DevMem *all[] = {&c->raw, &c->code, &c->seqof, &c->p2, &c->pm, &c->seq_off, &c->seq_set, &c->set_seq0, &c->set_base0,
                     &c->set_nmin, &c->seq_nmin, &c->dbl_off, &c->z0, &c->keysA, &c->keysB, &c->valsA, &c->valsB, &c->sa, &c->t0, &c->t1,
                     &c->t2, &c->t3, &c->t4, &c->t5, &c->shard_bounds, &c->bk_hist, &c->chb_sets, &c->chb_evbase, &c->chb_events, &c->chb_work, &c->chb_redo, &c->counter, &c->tiles, &c->pyr, &c->pyr2, &c->rs_start, &c->rs_count, &c->rs_cbase, &c->rs_stride, &c->sa0, &c->saidx0, &c->leaf_set, &c->lcp0, &c->set_nblocks,
                     &c->set_blk0, &c->set_pos0, &c->set_flags, &c->set_nchains, &c->set_cyclic, &c->firstmax, &c->set_collected, &c->set_suffixfree, &c->set_neff, &c->seq_per, &c->rare_collected, &c->rare_suffixfree, &c->blk_lb,
                     &c->blk_depth, &c->blk_set, &c->order, &c->o_depth, &c->o_set, &c->o_pos, &c->elem_blk, &c->seghead,
                     &c->succ_lo, &c->succ_hi, &c->next, &c->gap, &c->size, &c->total, &c->interval, &c->inv, &c->f_depth,
                     &c->f_size, &c->f_total, &c->f_interval, &c->f_next, &c->f_pos, &c->rotations, &c->sh_rec, &c->sh_sa0, &c->sh_saidx0, &c->sh_lcp0, &c->sh_allrec, &c->blk_leaf, &c->blk_tab, &c->f_leaf, &c->f_set, &c->let_off, &c->let_out};
    for (DevMem *m : all) dev_free(*m);
    for (int i = 0; i < 4; i++) dev_free(c->ps.block_sums[i]);
    dev_free(c->ps.counts);
    dev_free(c->ps.chain);
#ifndef CSA_EMU
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->meta_pinned) cudaFreeHost(c->meta_pinned);
    if (c->tm.ok) for (int i = 0; i < 8; i++) cudaEventDestroy(c->tm.ev[i]);
    cudaStreamDestroy(c->own_stream);
#else
    free(c->pinned);
    free(c->meta_pinned);
#endif
    delete c;
}

static int meta_reserve(csa_gpu_ctx *c, size_t bytes) { // page-locked staging of the batch's small tables
    if (c->meta_bytes >= bytes) return 0;
#ifndef CSA_EMU
    if (c->meta_pinned) cudaFreeHost(c->meta_pinned);
    c->meta_pinned = nullptr; c->meta_bytes = 0;
    CUDA_TRY(cudaMallocHost(&c->meta_pinned, bytes));
#else
    free(c->meta_pinned);
    c->meta_pinned = malloc(bytes);
    if (!c->meta_pinned) CSA_FAIL(CSA_GPU_ENOMEM, "out of host memory");
#endif
    c->meta_bytes = bytes;
    return 0;
}

static int pinned_reserve(csa_gpu_ctx *c, size_t bytes) {
    if (c->pinned_bytes >= bytes) return 0;
#ifndef CSA_EMU
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr; c->pinned_bytes = 0;
    CUDA_TRY(cudaMallocHost(&c->pinned, bytes));
#else
    free(c->pinned);
    c->pinned = malloc(bytes);
    if (!c->pinned) CSA_FAIL(CSA_GPU_ENOMEM, "out of host memory");
#endif
    c->pinned_bytes = bytes;
    return 0;
}

// ---- upload ----------------------------------------------------------------------------------------
// describe the batch; `fill` copies the letters of sequence k into the staging buffer
template <class Fill>
static int upload_common(csa_gpu_ctx *c, int nsets, const int *set_start, const long long *lens, Fill fill,
                         const char *direct = nullptr) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (nsets < 1 || !set_start) CSA_FAIL(CSA_GPU_EINVAL, "batch needs at least one set");
    if (set_start[0] != 0) CSA_FAIL(CSA_GPU_EINVAL, "set_start[0] must be 0");
    c->uploaded = false; c->ran = false;
    long long M = set_start[nsets];
    if (M < 2 || M >= (1ll << 31)) CSA_FAIL(CSA_GPU_EINVAL, "bad number of sequences %lld", M);
    c->nsets = nsets; c->M = (u32)M;
    c->h_seq_off.assign(M + 1, 0); c->h_seq_set.assign(M, 0);
    c->h_set_seq0.assign(nsets + 1, 0); c->h_set_base0.assign(nsets + 1, 0);
    c->h_set_nmin.assign(nsets, 0); c->h_z0.assign(nsets + 1, 0);
    c->h_dbl_off.assign(M + 1, 0);
    u64 tot = 0, dbl = 0;
    u32 nmax = 0, n0max = 0, mmax = 0, z = 0;
    for (int s = 0; s < nsets; s++) {
        int q0 = set_start[s], q1 = set_start[s + 1];
        if (q1 - q0 < 2) CSA_FAIL(CSA_GPU_EINVAL, "set %d has %d sequences; the path needs at least 2 (csamsa.c:533)", s, q1 - q0);
        c->h_set_seq0[s] = (u32)q0;
        c->h_set_base0[s] = (u32)tot;
        c->h_z0[s] = z;
        if ((u32)(q1 - q0) > mmax) mmax = (u32)(q1 - q0);
        u32 nmin = 0xFFFFFFFFu;
        for (int k = q0; k < q1; k++) {
            long long n = lens[k];
            if (n < 1) CSA_FAIL(CSA_GPU_EINVAL, "sequence %d is empty", k);
            if (tot + (u64)n >= (1ull << 31) - 64) CSA_FAIL(CSA_GPU_EINVAL, "batch larger than 2^31 bases; split it");
            c->h_seq_off[k] = (u32)tot;
            c->h_seq_set[k] = (u32)s;
            c->h_dbl_off[k] = dbl;
            tot += (u64)n;
            dbl += ((2 * (u64)n + 64 + 31) / 32) * 32;
            if ((u32)n > nmax) nmax = (u32)n;
            if ((u32)n < nmin) nmin = (u32)n;
            if (k == q0) { z += (u32)n; if ((u32)n > n0max) n0max = (u32)n; }
        }
        c->h_set_nmin[s] = nmin;
    }
    c->h_seq_off[M] = (u32)tot; c->h_dbl_off[M] = dbl;
    c->h_set_seq0[nsets] = (u32)M; c->h_set_base0[nsets] = (u32)tot; c->h_z0[nsets] = z;
    c->batch_nmin = *std::min_element(c->h_set_nmin.begin(), c->h_set_nmin.end());
    c->h_seq_nmin.resize(M);
    for (u32 k = 0; k < (u32)M; k++) c->h_seq_nmin[k] = c->h_set_nmin[c->h_seq_set[k]]; // (the same by sequence: one gather less in the word sort)
    c->max_set_bases = 0;
    for (int s = 0; s < nsets; s++) c->max_set_bases = std::max(c->max_set_bases, c->h_set_base0[s + 1] - c->h_set_base0[s]);
    c->N = (u32)tot; c->N0 = z; c->nmax = nmax; c->n0max = n0max; c->mmax = mmax; c->TW = dbl / 32 + 8; // (guard words: the word sort stages 5 words at a time and may read past the last sequence)
    u32 N = c->N;
    // stage the letters in pinned memory -- unless the caller's buffer is one contiguous, page-locked
    // block already (cudaHostRegister / csa_gpu_pin_host): then the copy engine reads it in place
    const void *src = direct;
#ifndef CSA_EMU
    if (src) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) != cudaSuccess || at.type != cudaMemoryTypeHost) { cudaGetLastError(); src = nullptr; }
    }
#endif
    if (!src) {
        TRY(pinned_reserve(c, (size_t)N));
        for (long long k = 0; k < M; k++) fill((int)k, (char *)c->pinned + c->h_seq_off[k]);
        src = c->pinned;
    }
#ifndef CSA_EMU
    CUDA_TRY(cudaSetDevice(c->device));
#endif
    Exec &ex = c->ex;
    TRY(dev_alloc(c->raw, N)); TRY(dev_alloc(c->code, N)); TRY(dev_alloc(c->seqof, sizeof(u32) * (size_t)N));
    TRY(dev_alloc(c->p2, sizeof(u64) * c->TW)); TRY(dev_alloc(c->pm, sizeof(u32) * c->TW));
    TRY(dev_alloc(c->seq_off, sizeof(u32) * (M + 1))); TRY(dev_alloc(c->seq_set, sizeof(u32) * M));
    TRY(dev_alloc(c->set_seq0, sizeof(u32) * (nsets + 1))); TRY(dev_alloc(c->set_base0, sizeof(u32) * (nsets + 1)));
    TRY(dev_alloc(c->set_nmin, sizeof(u32) * nsets)); TRY(dev_alloc(c->seq_nmin, sizeof(u32) * M)); TRY(dev_alloc(c->dbl_off, sizeof(u64) * (M + 1)));
    TRY(dev_alloc(c->z0, sizeof(u32) * (nsets + 1)));
    TRY(h2d(ex, c->raw.p, src, N));
    std::vector<u32> bs, bc, bb, bt; // tile-blocks of the first (segmented) sort: none straddles a set
    {
        u32 cbase = 0;
        for (int s = 0; s < nsets; s++) {
            u32 s0 = c->h_set_base0[s], s1 = c->h_set_base0[s + 1];
            u32 nb = (s1 - s0 + RS_TILE_ELEMS - 1) / RS_TILE_ELEMS;
            for (u32 j = 0; j < nb; j++) {
                bs.push_back(s0 + j * RS_TILE_ELEMS);
                bc.push_back(std::min<u32>(RS_TILE_ELEMS, s1 - (s0 + j * RS_TILE_ELEMS)));
                bb.push_back(cbase + j);
                bt.push_back(nb);
            }
            cbase += 256 * nb;
        }
        c->rs_nblocks = (u32)bs.size();
        size_t bytes = sizeof(u32) * bs.size();
        TRY(dev_alloc(c->rs_start, bytes)); TRY(dev_alloc(c->rs_count, bytes));
        TRY(dev_alloc(c->rs_cbase, bytes)); TRY(dev_alloc(c->rs_stride, bytes));
    }
    {   // the small tables: laid out one behind the other in page-locked memory, so that their copies are queued behind the
        // letters' copy without the host waiting for any of them (a copy from pageable memory returns only when it is staged)
        struct Tab { DevMem *dst; const void *src; size_t bytes; };
        const Tab tabs[] = {
            {&c->seq_off, c->h_seq_off.data(), sizeof(u32) * (M + 1)}, {&c->seq_set, c->h_seq_set.data(), sizeof(u32) * M},
            {&c->set_seq0, c->h_set_seq0.data(), sizeof(u32) * (nsets + 1)}, {&c->set_base0, c->h_set_base0.data(), sizeof(u32) * (nsets + 1)},
            {&c->set_nmin, c->h_set_nmin.data(), sizeof(u32) * nsets}, {&c->seq_nmin, c->h_seq_nmin.data(), sizeof(u32) * M},
            {&c->dbl_off, c->h_dbl_off.data(), sizeof(u64) * (M + 1)},
            {&c->z0, c->h_z0.data(), sizeof(u32) * (nsets + 1)},
            {&c->rs_start, bs.data(), sizeof(u32) * bs.size()}, {&c->rs_count, bc.data(), sizeof(u32) * bc.size()},
            {&c->rs_cbase, bb.data(), sizeof(u32) * bb.size()}, {&c->rs_stride, bt.data(), sizeof(u32) * bt.size()}};
        size_t total = 0;
        for (const Tab &t : tabs) total += (t.bytes + 15) & ~(size_t)15;
        TRY(meta_reserve(c, total));
        size_t at = 0;
        for (const Tab &t : tabs) {
            memcpy((char *)c->meta_pinned + at, t.src, t.bytes);
            TRY(h2d(ex, t.dst->p, (char *)c->meta_pinned + at, t.bytes));
            at += (t.bytes + 15) & ~(size_t)15;
        }
    }
    TRY(exec_sync(ex)); // the host vectors and the staging buffer may change after we return
    c->uploaded = true;
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_batch_upload(csa_gpu_ctx *c, int nsets, const int *set_start, const char *const *texts,
                                    const int *textsizes) {
    if (!texts || !textsizes || !set_start || nsets < 1) CSA_FAIL(CSA_GPU_EINVAL, "null argument");
    long long M = set_start[nsets];
    if (M < 0) CSA_FAIL(CSA_GPU_EINVAL, "bad set_start");
    std::vector<long long> lens((size_t)M);
    for (long long k = 0; k < M; k++) lens[k] = textsizes[k];
    return upload_common(c, nsets, set_start, lens.data(),
                         [&](int k, char *dst) { memcpy(dst, texts[k], (size_t)textsizes[k]); });
}

extern "C" int csa_gpu_batch_upload_flat(csa_gpu_ctx *c, int nsets, const int *set_start, const char *text,
                                         const long long *text_start) {
    if (!text || !text_start || !set_start || nsets < 1) CSA_FAIL(CSA_GPU_EINVAL, "null argument");
    long long M = set_start[nsets];
    if (M < 0) CSA_FAIL(CSA_GPU_EINVAL, "bad set_start");
    std::vector<long long> lens((size_t)M);
    for (long long k = 0; k < M; k++) lens[k] = text_start[k + 1] - text_start[k];
    return upload_common(c, nsets, set_start, lens.data(),
                         [&](int k, char *dst) { memcpy(dst, text + text_start[k], (size_t)(text_start[k + 1] - text_start[k])); },
                         text + text_start[0]);
}

// page-lock a host buffer so that uploads from it need no staging copy (cudaHostRegister)
extern "C" int csa_gpu_pin_host(const void *p, unsigned long long bytes) {
#ifndef CSA_EMU
    CUDA_TRY(cudaHostRegister(const_cast<void *>(p), (size_t)bytes, cudaHostRegisterDefault));
#else
    (void)p; (void)bytes;
#endif
    return CSA_GPU_OK;
}
extern "C" int csa_gpu_unpin_host(const void *p) {
#ifndef CSA_EMU
    CUDA_TRY(cudaHostUnregister(const_cast<void *>(p)));
#else
    (void)p;
#endif
    return CSA_GPU_OK;
}

// ---- run --------------------------------------------------------------------------------------------
static inline void mark(csa_gpu_ctx *c, int i) {
#ifndef CSA_EMU
    cudaEventRecord(c->tm.ev[i], c->ex.stream);
#else
    (void)c; (void)i;
#endif
}

static int read_u32(csa_gpu_ctx *c, const void *dev, u32 *out) { return d2h(c->ex, out, dev, sizeof(u32)); }

// sort helper on the context's double buffers: on return keysA/valsA hold the sorted data
static int sort_pairs(csa_gpu_ctx *c, long long n, int begin_bit, int end_bit) {
    u64 *k = P<u64>(c->keysA), *ka = P<u64>(c->keysB);
    u32 *v = P<u32>(c->valsA), *va = P<u32>(c->valsB);
    TRY(radix_sort_pairs<u64>(c->ex, c->ps, k, v, ka, va, n, begin_bit, end_bit));
    if (k != P<u64>(c->keysA)) { std::swap(c->keysA, c->keysB); std::swap(c->valsA, c->valsB); }
    return 0;
}

// head[]/rank[] from the sorted keys in keysA (device-wide path)
// (first sort: lcp != nullptr also gives every border its LCP, read off the keys; rank == nullptr: heads only)
static int heads_and_ranks(csa_gpu_ctx *c, u32 *head, u32 *rank, u32 *counter, u32 *ngroups, bool keys32 = false,
                           bool fix_set_starts = false, u32 *lcp = nullptr, int letters = 0, int lbits = 0) {
    Exec &ex = c->ex;
    u32 N = c->N;
    TRY(dev_zero(ex, counter, sizeof(u32)));
    { FlagArgs a{P<u64>(c->keysA), keys32 ? P<u32>(c->keysA) : nullptr, head, counter, lcp, letters, lbits,
                 P<u32>(c->valsA), P<u32>(c->seqof), P<u32>(c->seq_off), c->batch_nmin < (u32)letters ? 1 : 0, 0u}; launch_flag(ex, N, a); }
    if (fix_set_starts) { SetStartArgs a{view_of(c), head, counter, lcp}; launch_setstart(ex, c->nsets, a); }
    TRY((scan_u32<ScanMax, true>(ex, c->ps, head, head, N)));
    if (rank) { SetRankArgs a{P<u32>(c->valsA), head, rank}; launch_setrank(ex, N, a); }
    return read_u32(c, counter, ngroups);
}

// phase 0: the whole stage.  Sharded over the ranks of a job (csa_gpu_shard_*): phase 1 = first sort (every rank
// the same), bucket borders, word sort of this rank's bucket; phase 2 = what is left after the buckets were exchanged.
static int stage_suffix_array(csa_gpu_ctx *c, const BatchView &v, int phase = 0) {
    Exec &ex = c->ex;
    u32 N = c->N;
    u32 *head = P<u32>(c->t0), *rank = P<u32>(c->t1), *rank2 = P<u32>(c->t3), *counter = P<u32>(c->counter);
    u32 ntiles = (N + RF_NOMINAL - 1) / RF_NOMINAL;
    TRY(dev_alloc(c->tiles, sizeof(u32) * ((size_t)ntiles + 2)));
    u32 any_other = c->sa_any_other;
    if (phase != 2) TRY(read_u32(c, counter + 8, &any_other)); // set by k_encode: a letter outside ACGT somewhere in the batch
    c->sa_any_other = any_other;
    // 12 letters tell the places of a 16 kb mitogenome apart; in a 5 Mb chromosome every fourth 12-letter word turns up again
    // somewhere else (4^12 = 16.7 M) and drags a whole unrelated column into the group: 16 letters there (one more radix pass)
    static const int env_letters = getenv("CSA_GPU_KEY_LETTERS") ? atoi(getenv("CSA_GPU_KEY_LETTERS")) : 0;
    const int letters32 = env_letters == 12 || env_letters == 16 ? env_letters : (c->max_set_bases > (1u << 20) ? 16 : 12);
    const int letters = any_other ? CSA_K0 : letters32, lbits = any_other ? CSA_LETTER_BITS : 2;
    int nbits = bits_for((u64)N - 1);
    u64 sorted_len = (u64)letters;
    u32 ngroups = c->sa_ngroups, maxg = 0;
    bool words = true;
    u32 *lcp = P<u32>(c->t5);
    // a job of several ranks on ONE set: every rank sorts only its own bucket (ranges of the key's first 6 letters)
    const bool own_sort = phase == 1 && c->nsets == 1 && c->shard_nranks > 1 && !c->shard_full_sort;
    if (phase == 1) c->shard_own_sort = own_sort;
    if (own_sort) {
        const int R = c->shard_nranks, kbits = letters * lbits, shift = kbits - BK_BITS;
        { InitKeyArgs a{v, any_other ? P<u64>(c->keysB) : nullptr, any_other ? nullptr : P<u32>(c->keysB), P<u32>(c->valsB), letters32};
          launch_initkey(ex, N, a); }
        TRY(dev_alloc(c->bk_hist, sizeof(u32) * BK_BINS));
        TRY(dev_zero(ex, c->bk_hist.p, sizeof(u32) * BK_BINS));
        BucketArgs b{any_other ? P<u64>(c->keysB) : nullptr, any_other ? nullptr : P<u32>(c->keysB), P<u32>(c->valsB), N, shift,
                     P<u32>(c->bk_hist), 0u, 0u, P<u32>(c->t0), P<u32>(c->t1), nullptr, nullptr, nullptr};
        launch_bkhist(ex, N, b);
        std::vector<u32> h(BK_BINS), cut(R + 1, (u32)BK_BINS);
        TRY(d2h(ex, h.data(), c->bk_hist.p, sizeof(u32) * BK_BINS));
        c->h_shard_bounds.assign(R + 1, N);
        cut[0] = 0; c->h_shard_bounds[0] = 0;
        {   // bucket r starts at the first prefix whose suffixes begin at or after r * N / R
            u64 cum = 0;
            int r = 1;
            for (u32 p = 0; p < (u32)BK_BINS && r < R; p++) {
                while (r < R && cum >= (u64)r * N / R) { cut[r] = p; c->h_shard_bounds[r] = (u32)cum; r++; }
                cum += h[p];
            }
        }
        const u32 off = c->h_shard_bounds[c->shard_rank], n_r = c->h_shard_bounds[c->shard_rank + 1] - off;
        b.plo = cut[c->shard_rank]; b.phi = cut[c->shard_rank + 1];
        b.out64 = P<u64>(c->keysA) + off; b.out32 = P<u32>(c->keysA) + off; b.outv = P<u32>(c->valsA) + off;
        launch_bkflag(ex, N, b);
        TRY((scan_u32<ScanSum, false>(ex, c->ps, P<u32>(c->t0), P<u32>(c->t1), N)));
        launch_bkscatter(ex, N, b);
        if (any_other) {
            u64 *k = P<u64>(c->keysA) + off, *ka = P<u64>(c->keysB) + off;
            u32 *vv = P<u32>(c->valsA) + off, *va = P<u32>(c->valsB) + off;
            TRY(radix_sort_pairs<u64>(ex, c->ps, k, vv, ka, va, n_r, 0, kbits));
            if (vv != P<u32>(c->valsA) + off) { std::swap(c->keysA, c->keysB); std::swap(c->valsA, c->valsB); }
        } else {
            u32 *k = P<u32>(c->keysA) + off, *ka = P<u32>(c->keysB) + off;
            u32 *vv = P<u32>(c->valsA) + off, *va = P<u32>(c->valsB) + off;
            TRY(radix_sort_pairs<u32>(ex, c->ps, k, vv, ka, va, n_r, 0, kbits));
            if (vv != P<u32>(c->valsA) + off) { std::swap(c->keysA, c->keysB); std::swap(c->valsA, c->valsB); }
        }
        TRY(dev_fill_ff(ex, lcp, sizeof(u32) * (size_t)N));
        TRY(dev_zero(ex, counter, sizeof(u32)));
        { FlagArgs a{P<u64>(c->keysA) + off, any_other ? nullptr : P<u32>(c->keysA) + off, head + off, counter, lcp + off, letters, lbits,
                     P<u32>(c->valsA) + off, P<u32>(c->seqof), P<u32>(c->seq_off), c->batch_nmin < (u32)letters ? 1 : 0, off};
          launch_flag(ex, n_r, a); }
        TRY((scan_u32<ScanMax, true>(ex, c->ps, head + off, head + off, n_r)));
        ngroups = 0; // (not counted: the word sort looks at every group of the bucket anyway)
        c->sa_ngroups = 0;
        c->lcp_state = 0; c->ws_runs = 0;
        c->rounds_tiled = c->rounds_global = c->rounds_quad = c->rounds_list = 0;
    } else if (phase != 2) {
    RsSeg seg{P<u32>(c->rs_start), P<u32>(c->rs_count), P<u32>(c->rs_cbase), P<u32>(c->rs_stride), c->rs_nblocks,
              P<u32>(c->set_base0), (u32)c->nsets};
    { InitKeyArgs a{v, any_other ? P<u64>(c->keysA) : nullptr, any_other ? nullptr : P<u32>(c->keysA), nullptr, letters32};
      launch_initkey(ex, N, a); }
    if (any_other) {
        u64 *k = P<u64>(c->keysA), *ka = P<u64>(c->keysB);
        u32 *vv = P<u32>(c->valsA), *va = P<u32>(c->valsB);
        TRY(radix_sort_pairs<u64>(ex, c->ps, k, vv, ka, va, N, 0, letters * lbits, &seg, true));
        if (vv != P<u32>(c->valsA)) { std::swap(c->keysA, c->keysB); std::swap(c->valsA, c->valsB); }
    } else {
        u32 *k = P<u32>(c->keysA), *ka = P<u32>(c->keysB);
        u32 *vv = P<u32>(c->valsA), *va = P<u32>(c->valsB);
        TRY(radix_sort_pairs<u32>(ex, c->ps, k, vv, ka, va, N, 0, letters * lbits, &seg, true));
        if (vv != P<u32>(c->valsA)) { std::swap(c->keysA, c->keysB); std::swap(c->valsA, c->valsB); }
    }
    ngroups = 0;
    // word sort (k_wsort) first; it leaves what it cannot finish (sets of very short sequences, long repeats)
    // It compares every pair of suffixes of a group, so it is the choice when groups are small (a handful of
    // related genomes per set): under WS_PAIRS_PER_SUFFIX pairs per suffix of the batch.  Larger groups (dozens of
    // near-identical sequences) are cheaper by rank doubling, which never reads a letter twice.
    words = phase == 1 || c->round_mode == 0 || c->round_mode == 4;
    c->lcp_state = 0;
    c->ws_runs = 0;
    if (words) TRY(dev_fill_ff(ex, lcp, sizeof(u32) * (size_t)N));
    // group borders, heads, and what decides between the word sort and the doubling rounds -- the number of groups, the
    // largest one, the pairs they hold -- all queued, then ONE wait for the four numbers
    unsigned long long *pairs = (unsigned long long *)(counter + 22), hpairs[4] = {0, 0, 0, 0};
    c->carry_pick = false;
    c->carry_big = false;
    TRY(dev_zero(ex, counter, 4 * sizeof(u32)));
    TRY(dev_zero(ex, pairs, 4 * sizeof(*pairs)));
    { FlagArgs a{P<u64>(c->keysA), !any_other ? P<u32>(c->keysA) : nullptr, head, counter, words ? lcp : nullptr, letters, lbits,
                 P<u32>(c->valsA), P<u32>(c->seqof), P<u32>(c->seq_off), c->batch_nmin < (u32)letters ? 1 : 0, 0u}; launch_flag(ex, N, a); }
    { SetStartArgs a{view_of(c), head, counter, words ? lcp : nullptr}; launch_setstart(ex, c->nsets, a); }
    TRY((scan_u32<ScanMax, true>(ex, c->ps, head, head, N)));
    { MaxGroupArgs a{head, counter + 2, N, pairs}; launch_maxgroup(ex, N, a); }
    {
        u32 hc[30];
        TRY(d2h(ex, hc, counter, sizeof(hc))); // [0] groups, [2] largest group, [22..29] pairs, suffixes that share a group, ... of no more than 32, ... than 256
        ngroups = hc[0]; maxg = hc[2];
        memcpy(hpairs, hc + 22, sizeof(hpairs));
    }
    c->rounds_tiled = c->rounds_global = c->rounds_quad = c->rounds_list = 0;
    if (ngroups != N) {
        c->ws_pairs = (double)hpairs[0];
        c->ws_sharing = (double)hpairs[1];
        if (getenv("CSA_GPU_TRACE")) fprintf(stderr, "[csa] groups %u of %u suffixes, largest %u, pairs %.0f (%.2f per suffix), sharing %.0f\n", ngroups, N, maxg, c->ws_pairs, c->ws_pairs / N, c->ws_sharing);
        // measured: ~0.05 ns per pair; rank doubling + LCP ~0.2 ns per suffix while a set's ranks live in L2, ~0.55 ns when
        // they do not (one set of tens of millions of suffixes: every gather a trip to HBM)
        const double per_suffix = c->max_set_bases > WS_LARGE_SET ? WS_PAIRS_PER_SUFFIX_LARGE : WS_PAIRS_PER_SUFFIX;
        if (phase == 0 && c->round_mode == 0 && !c->ws_force && (double)hpairs[0] > per_suffix * (double)N) {
            // many pairs = many near-identical sequences: every pair compared is too much, but when the groups fit a warp
            // (dozens of sequences, not hundreds) a column's order carries over to the next and few pairs are compared at all
            // (hundreds: a CTA per group walks)
            if (c->carry_mode == 0 && (double)hpairs[2] >= 0.9 * (double)hpairs[1]) c->carry_pick = true;
            else if (c->carry_mode == 0 && c->max_set_bases <= WS_LARGE_SET && (double)hpairs[3] >= 0.9 * (double)hpairs[1]) c->carry_pick = c->carry_big = true;
            else words = false;
        }
        // (forced, tests: whatever the largest group asks for)
        if (phase == 0 && c->carry_mode == 1 && c->max_set_bases <= WS_LARGE_SET && maxg > CY_MAXG) c->carry_big = true;
    }
    if (!words) { SetRankArgs r{P<u32>(c->valsA), head, rank}; launch_setrank(ex, N, r); }
    c->sa_ngroups = ngroups;
    } // phase != 2
    // counter[0] groups, [1] a tile would overflow, [2] largest group
    auto largest_group = [&](u32 *out) -> int {
        TRY(dev_zero(ex, counter + 2, sizeof(u32)));
        { MaxGroupArgs a{head, counter + 2, N, nullptr}; launch_maxgroup(ex, N, a); }
        return read_u32(c, counter + 2, out);
    };
    // gencycsuffixtrees.c compares rotations letter by letter; two periodic strings that agree on
    // n_a+n_b letters agree for ever, so 2*nmax sorted letters settle every comparison
    // group lists (see k_refine_g): two lists of (start:size), two of fresh singletons, ping-ponged.
    // They live in buffers that are idle during this stage; a device-wide round clobbers the scratch.
    u64 *glist[2] = {P<u64>(c->t4), P<u64>(c->t2)};      // N/2 entries of 8 B each fit 4 B x N
    bool list_valid = false;
    u32 nlist = 0, nsingles = 0, active_est = N;
    int cur = 0;
    auto singles_buf = [&](int which) { return which == 0 ? P<u32>(c->sa) : P<u32>(c->valsB); };
    auto flush_singles = [&]() -> int { // ranks of last round's new singletons into the other rank buffer
        if (nsingles) { CopySinglesArgs a{singles_buf(cur), rank, rank2, head}; launch_copysingles(ex, nsingles, a); }
        nsingles = 0;
        return 0;
    };
    if (words) {
        c->lcp_state = 1;
        if (phase == 1 && !own_sort) { // this rank's bucket: borders at group borders, the same on every rank
            BoundsArgs b{head, N, (u32)c->shard_nranks, P<u32>(c->shard_bounds)};
            launch_bounds(ex, b);
            c->h_shard_bounds.assign(c->shard_nranks + 1, 0);
            TRY(d2h(ex, c->h_shard_bounds.data(), c->shard_bounds.p, sizeof(u32) * (c->shard_nranks + 1)));
        }
        if (ngroups != N && phase != 2) {
            u32 *res = counter + 16;
            const u32 init[6] = {0u, 0u, 0xFFFFFFFFu, 0u, 0u, 0u};
            TRY(h2d(ex, res, init, sizeof(init)));
            const u32 lo = phase == 1 ? c->h_shard_bounds[c->shard_rank] : 0u, hi = phase == 1 ? c->h_shard_bounds[c->shard_rank + 1] : N;
            WSortArgs a{v, P<u32>(c->valsA), head, lcp, N, lo, hi, (u32)letters, c->ws_depth_cap, (int)any_other, glist[0], glist[1], 0u, res,
                        nullptr, nullptr, 0u, nullptr, nullptr, CY_MAXG, nullptr, nullptr, 0u};
            // sets of whole genomes (millions of letters a sequence): what near-identical genomes share runs for hundreds of
            // letters, and a column's order carries over to the next (pipeline.cuh "carried word sort")
            const bool carry = c->carry_mode == 1 || (c->carry_mode == 0 && (c->max_set_bases > WS_LARGE_SET || c->carry_pick));
            c->ws_carried = carry;
            if (carry) {
                u32 *grp = rank, *head2 = rank2, *roots = P<u32>(c->keysA), *nroots = counter + 32;
                const bool cybig = carry && c->carry_big && phase == 0; // groups of up to CY_BIGG walked (a CTA each)
                u32 *roots2 = roots + (size_t)N / 2 + 1, *nroots2 = counter + 38; // (at most N / 2 roots, N / 33 of them long ones)
                unsigned char *flag = P<unsigned char>(c->keysB);
                if (phase == 1) TRY(dev_fill_ff(ex, grp, sizeof(u32) * (size_t)N)); // (suffixes of other ranks' buckets: in no group)
                TRY(dev_zero(ex, nroots, 2 * sizeof(u32)));
                TRY(dev_zero(ex, nroots2, 2 * sizeof(u32)));
                static const bool direct = getenv("CSA_GPU_CYGRP_DIRECT") != nullptr; // (experiments)
                const bool dealt = !direct && (c->carry_mode == 1 || c->max_set_bases > WS_LARGE_SET) && hi - lo > 1; // (sets of a few MB: their stretch of grp sits in L2 anyway)
                const int pack = (N < (1u << 27) && !c->cy_nopack && !cybig) ? 1 : 0;
                u32 *startbits = P<u32>(c->keysB) + ((size_t)N + 3) / 4 + 4; // (behind the flags: N bits)
                TRY(dev_zero(ex, startbits, sizeof(u32) * ((size_t)N / 32 + 1)));
                { CySeqBitsArgs sb{v, startbits}; launch_cyseqbits(ex, (long long)c->M, sb); }
                CarryArgs ca{v, P<u32>(c->valsA), head, head2, grp, flag, lo, hi, pack, dealt ? P<u32>(c->valsB) : nullptr, cybig ? CY_BIGG : CY_MAXG, startbits};
                launch_cygrp(ex, (long long)hi - lo, ca);
                if (dealt) {
                    u32 *k = P<u32>(c->valsA) + lo, *ka = P<u32>(c->sa), *vv = P<u32>(c->valsB), *va = P<u32>(c->t2);
                    TRY(radix_sort_pairs<u32>(ex, c->ps, k, vv, ka, va, (long long)hi - lo, nbits > 8 ? nbits - 8 : 0, nbits));
                    CyScatterArgs sc{k, vv, grp};
                    launch_cyscatter(ex, (long long)hi - lo, sc);
                    ca.gval = nullptr;
                }
                launch_cyroots(ex, (long long)hi - lo, ca);
                u64 *list = P<u64>(c->keysA) + ((size_t)N + 1) / 2; // (behind the roots; at most N / 2 groups of two or more)
                u32 *nlistA = counter + 34, *nlistB = counter + 36; // (each followed by its work counter)
                TRY(dev_zero(ex, nlistA, 4 * sizeof(u32)));
                { CyListArgs l{head2, flag, 1u, lo, hi, list, nlistA}; launch_cylist(ex, (long long)hi - lo, l); }
                a.head_in = head2; a.flag = flag; a.want = 1u; a.roots = roots; a.nroots = nroots;
                a.maxg = ca.maxg; a.roots2 = roots2; a.nroots2 = nroots2;
                if (cybig && !any_other) a.list_cap = CY_MAXG; // (ACGT only: longer roots word by word, k_wsort_words; measured with 64 and 128: slower)
                a.words_small = any_other ? 0u : 1u;           // (... and the shorter ones by a warp, ws_words_warp)
                launch_wsort_list(ex, a, list, nlistA);
                if (cybig) { // roots longer than a warp's window: one CTA each, BEFORE the walks (they start from these too)
                    TRY(d2h(ex, c->ws_left, res, sizeof(c->ws_left)));
                    if (c->ws_left[5]) {
                        a.nbig = c->ws_left[5];
                        launch_wsort_big(ex, a);
                        a.nbig = 0;
                        TRY(dev_zero(ex, res + 5, sizeof(u32)));
                    }
                }
                CyWalkArgs w{v, P<u32>(c->valsA), head, lcp, head2, grp, pack, flag, lo, hi, (u32)letters, roots, nroots, nroots + 1,
                             roots2, nroots2, nroots2 + 1};
                launch_cywalk(ex, w);
                if (cybig) launch_cywalk_cta(ex, w);
                a.want = 0u; a.roots = nullptr; a.nroots = nullptr; a.list_cap = 0u;
                { CyListArgs l{head2, flag, 0u, lo, hi, list, nlistB}; launch_cylist(ex, (long long)hi - lo, l); }
                launch_wsort_list(ex, a, list, nlistB);
                if (getenv("CSA_GPU_TRACE")) {
                    std::vector<unsigned char> f((size_t)hi - lo);
                    std::vector<u32> h2((size_t)hi - lo);
                    u32 nr = 0, nr2 = 0;
                    TRY(d2h(ex, f.data(), flag + lo, f.size())); TRY(d2h(ex, h2.data(), head2 + lo, sizeof(u32) * h2.size()));
                    TRY(d2h(ex, &nr, nroots, sizeof(u32))); TRY(d2h(ex, &nr2, nroots2, sizeof(u32)));
                    size_t cnt[3] = {0, 0, 0};
                    for (size_t x = 0; x + 1 < f.size(); x++) if (h2[x] == lo + x && h2[x + 1] == lo + x) cnt[f[x] < 3 ? f[x] : 0]++;
                    fprintf(stderr, "[csa] carried word sort: %zu groups ordered by letters (%u + %u walks started), %zu written by the walks, %zu left to the sweep\n",
                            cnt[1], nr, nr2, cnt[2], cnt[0]);
                }
            } else launch_wsort(ex, a);
            TRY(d2h(ex, c->ws_left, res, sizeof(c->ws_left)));
            if (c->ws_left[5]) { // groups that did not fit a warp's window: one CTA each
                a.nbig = c->ws_left[5];
                launch_wsort_big(ex, a);
                TRY(d2h(ex, c->ws_left, res, sizeof(c->ws_left)));
#ifndef CSA_EMU
                if (ex.prof && !ex.prof->recs.empty()) ex.prof->recs.back().bytes = 8.0 * a.nbig;
#endif
            }
#ifndef CSA_EMU
            // every suffix's head in; per suffix of a group: sa in, sa + head + lcp out
            // (suffixes of a group: counted by k_maxgroup; a bucket's share of them when the stage is sharded)
            if (ex.prof) for (auto &r : ex.prof->recs) {
                if (!strcmp(r.name, "k_wsort")) r.bytes = (4.0 * N + 16.0 * c->ws_sharing) * ((double)(hi - lo) / N);
                // the walks: per suffix of a group its place in the group table in, sa + head + lcp out (the few groups ordered by
                // letters are not told apart here: no count of them comes back to the host)
                if (!strcmp(r.name, "k_cywalk") && !c->carry_big) r.bytes = 16.0 * (phase == 1 ? (double)(hi - lo) : c->ws_sharing);
                if (!strcmp(r.name, "k_cywalk_cta")) r.bytes = 16.0 * c->ws_sharing; // (the warps' walks' share is not told apart)
            }
#endif
            c->ws_runs = 1;
        }
        if (phase == 1) return 0; // the caller exchanges the buckets and the lists of what is left, then phase 2
        if (ngroups != N) {
            if (c->ws_left[0] == 0) ngroups = N;
            else { // some groups run deeper than the walk went (or were too long for a warp): doubling rounds from there
                { SetRankArgs r{P<u32>(c->valsA), head, rank}; launch_setrank(ex, N, r); }
                TRY(d2d(ex, rank2, rank, sizeof(u32) * (size_t)N));
                nlist = c->ws_left[0]; active_est = c->ws_left[1]; sorted_len = c->ws_left[2]; maxg = c->ws_left[3];
                ngroups = N - active_est + nlist;
                list_valid = true;
                c->lcp_state = active_est <= N / 8 ? 2 : 0;
            }
        }
        if (phase == 2 && c->shard_own_sort) { // what two neighbouring buckets share at their border: no rank has seen both keys
            for (int r = 1; r < c->shard_nranks; r++)
                if (c->h_shard_bounds[r] > 0 && c->h_shard_bounds[r] < N) TRY(dev_fill_ff(ex, lcp + c->h_shard_bounds[r], sizeof(u32)));
            if (c->lcp_state == 1) c->lcp_state = 2;
        }
    }
    while (ngroups != N && sorted_len < 2ull * c->nmax) {
        // lists pay off once most suffixes stand alone; while nearly all still share a group the tile
        // rounds stream them at the same cost without the list upkeep
        // (and while the groups are not tiny: one warp per group wastes its lanes on pairs and triples)
        const u32 sharing_groups = ngroups > N - active_est ? ngroups - (N - active_est) : 1;
        const bool want_list = (c->round_mode == 0 || c->round_mode >= 4) && maxg <= RF_QUAD_GROUP &&
                               (list_valid || (active_est < N / 2 && active_est / sharing_groups >= 6));
        if (want_list) {
            if (!list_valid) {
                TRY(dev_zero(ex, counter + 4, sizeof(u32)));
                { GListBuildArgs a{head, N, glist[cur], counter + 4}; launch_glist_build(ex, a); }
                TRY(read_u32(c, counter + 4, &nlist));
                // from here on only listed suffixes are touched: both rank buffers must agree on all others
                TRY(d2d(ex, rank2, rank, sizeof(u32) * (size_t)N));
                nsingles = 0;
                list_valid = true;
            }
            TRY(flush_singles());
            const int nkeys = (4 * sorted_len < (1ull << 31)) ? 3 : 1;
            TRY(dev_zero(ex, counter + 2, 4 * sizeof(u32))); // [2] largest group [3] staged [4] next list [5] new singletons
            RefineGArgs a{v, P<u32>(c->valsA), head, rank, rank2, (u32)sorted_len, nkeys, glist[cur], nlist,
                          glist[cur ^ 1], counter + 4, singles_buf(cur ^ 1), counter + 5, counter + 2, counter + 3,
                          P<u32>(c->keysA), P<u32>(c->keysA) + N, P<u32>(c->keysB), P<u32>(c->keysB) + N};
            launch_refine_g(ex, a);
            std::swap(rank, rank2);
            cur ^= 1;
            u32 res[4];
            TRY(d2h(ex, res, counter + 2, sizeof(res)));
#ifndef CSA_EMU
            if (ex.prof && !ex.prof->recs.empty()) ex.prof->recs.back().bytes = (nkeys == 3 ? 36.0 : 28.0) * res[1] + 8.0 * nlist;
#endif
            maxg = res[0] > 1 ? res[0] : 1; nlist = res[2]; nsingles = res[3]; active_est = res[1];
            if (nlist == 0) ngroups = N; // nothing shares a group any more
            else ngroups = (N - res[1]) + nlist + nsingles; // (rough while lists run; exact again after a tile round)
            c->rounds_list++;
            sorted_len *= (nkeys == 3 ? 4 : 2);
            continue;
        }
        if (list_valid) { TRY(flush_singles()); list_valid = false; }
        TRY(dev_zero(ex, counter, 5 * sizeof(u32)));
        { TileArgs a{head, P<u32>(c->tiles), counter + 1, N, ntiles}; launch_tile(ex, ntiles, a); }
        u32 oversize = 0;
        if (maxg > RF_NOMINAL) TRY(read_u32(c, counter + 1, &oversize)); // smaller groups always fit a tile
        if (!oversize && !c->force_global_rounds) {
            RefineArgs a{v, P<u32>(c->valsA), head, rank, rank2, P<u32>(c->tiles), (u32)sorted_len, counter, ntiles, counter + 2, counter + 3, counter + 4};
            // small groups: four times the letters per round (three rank gathers); else twice
            const bool quad = maxg <= RF_QUAD_GROUP && !c->no_quad_rounds && 4 * sorted_len < (1ull << 31);
            if (quad) launch_refine4(ex, a); else launch_refine(ex, a);
            std::swap(rank, rank2);
            u32 res[5];
            TRY(d2h(ex, res, counter, sizeof(res)));
            ngroups = res[0]; maxg = res[2]; active_est = res[4];
#ifndef CSA_EMU
            // the profile counts what the round really had to move: every suffix's head (4 B), and for
            // the res[3] suffixes not settled yet the rest (sa in/out, sequence, ranks gathered, head and rank out)
            if (ex.prof && !ex.prof->recs.empty()) ex.prof->recs.back().bytes = 4.0 * N + (quad ? 32.0 : 24.0) * res[3];
#endif
            if (quad) { c->rounds_quad++; sorted_len *= 4; } else { c->rounds_tiled++; sorted_len *= 2; }
        } else { // a group larger than a tile: device-wide radix sort of (rank, rank h letters on)
            { Key2Args a{v, P<u32>(c->valsA), rank, P<u64>(c->keysA), (u32)sorted_len, nbits}; launch_key2(ex, N, a); }
            TRY(sort_pairs(c, N, 0, 2 * nbits));
            TRY(heads_and_ranks(c, head, rank, counter, &ngroups));
            if (ngroups != N) TRY(largest_group(&maxg));
            c->rounds_global++;
            sorted_len *= 2;
        }
    }
    std::swap(c->sa, c->valsA); // (both N x u32: the sorted suffixes become `sa`, the old buffer the next sort's scratch)
    return 0;
}

// block candidates (collectNodes / removeSuffixNodes / removeNonUniqueNodes on ordinary sets): through the cover array R[]
// when the counts are asked for or a set holds more than 256 sequences, else straight from the LCP array (k_blockfind2)
static int stage_common_blocks(csa_gpu_ctx *c, const BatchView &v) {
    Exec &ex = c->ex;
    u32 N = c->N;
    int nsets = c->nsets;
    u32 *sa = P<u32>(c->sa), *lcp = P<u32>(c->t5), *nxt = P<u32>(c->t0), *R = P<u32>(c->t1);
    u32 *isblock = P<u32>(c->t0), *depth = P<u32>(c->t3);
    TRY(dev_alloc(c->saidx0, sizeof(u32) * (size_t)c->N0));
    if (!c->use_cover) {
        { BlockFind2Args a{v, sa, lcp, isblock, depth, c->mmax, 0u}; launch_blockfind2(ex, N, a); }
        return 0;
    }
    TRY(dev_zero(ex, c->firstmax.p, sizeof(u32) * nsets));
    { ColorKeyArgs a{v, sa, P<u64>(c->keysA), P<u32>(c->valsA)}; launch_colorkey(ex, N, a); }
    TRY(sort_pairs(c, N, 0, bits_for((u64)c->mmax - 1)));
    // colour 0 first: the SA places of sequence 0 of every set, kept for the block-order stage (k_seq0take)
    TRY(d2d(ex, c->saidx0.p, c->valsA.p, sizeof(u32) * (size_t)c->N0));
    { NextArgs a{v, P<u64>(c->keysA), P<u32>(c->valsA), nxt, P<u32>(c->firstmax)}; launch_next(ex, N, a); }
    { CoverArgs a{v, sa, nxt, P<u32>(c->firstmax), R}; launch_cover(ex, N, a); launch_coverstart(ex, nsets, a); }
    TRY((scan_u32<ScanMax, true>(ex, c->ps, R, R, N)));
    { BlockFindArgs a{v, sa, lcp, R, isblock, depth, c->mmax}; launch_blockfind(ex, N, a); }
    return 0;
}

// the block borders compacted into the batch's block list (after rare.cuh had its say on marked sets)
static int stage_emit_blocks(csa_gpu_ctx *c, const BatchView &v) {
    Exec &ex = c->ex;
    u32 N = c->N;
    int nsets = c->nsets;
    u32 *sa = P<u32>(c->sa);
    u32 *isblock = P<u32>(c->t0), *depth = P<u32>(c->t3), *bidx = P<u32>(c->t4);
    TRY((scan_u32<ScanSum, false>(ex, c->ps, isblock, bidx, N)));
    { SetCountArgs a{v, isblock, bidx, P<u32>(c->set_nblocks)}; launch_setcount(ex, nsets, a); }
    c->h_set_nblocks.assign(nsets, 0);
    TRY(d2h(ex, c->h_set_nblocks.data(), c->set_nblocks.p, sizeof(u32) * nsets)); // (the one wait of this stage)
    u64 btot = 0;
    for (u32 x : c->h_set_nblocks) btot += x;
    c->B = (u32)btot;
    u32 B = c->B;
    TRY(dev_alloc(c->blk_lb, sizeof(u32) * (size_t)B)); TRY(dev_alloc(c->blk_depth, sizeof(u32) * (size_t)B));
    TRY(dev_alloc(c->blk_set, sizeof(u32) * (size_t)B));
    { BlockEmitArgs a{v, sa, isblock, bidx, depth, P<u32>(c->blk_lb), P<u32>(c->blk_depth), P<u32>(c->blk_set), P<u32>(c->set_nblocks)};
      launch_blockemit(ex, N, a); }
    c->h_set_blk0.assign(nsets + 1, 0); c->h_set_pos0.assign(nsets + 1, 0);
    u64 e = 0;
    u32 b = 0;
    for (int s = 0; s < nsets; s++) {
        c->h_set_blk0[s] = b; c->h_set_pos0[s] = (u32)e;
        b += c->h_set_nblocks[s];
        e += (u64)c->h_set_nblocks[s] * (c->h_set_seq0[s + 1] - c->h_set_seq0[s]);
    }
    c->h_set_blk0[nsets] = b; c->h_set_pos0[nsets] = (u32)e;
    c->E = (u32)e;
    TRY(h2d(ex, c->set_blk0.p, c->h_set_blk0.data(), sizeof(u32) * (nsets + 1)));
    TRY(h2d(ex, c->set_pos0.p, c->h_set_pos0.data(), sizeof(u32) * (nsets + 1)));
    return 0;
}

// csamsa.c:332,338: how many nodes collectNodes finds and removeSuffixNodes leaves (optional)
static int stage_stats(csa_gpu_ctx *c, const BatchView &v) {
    Exec &ex = c->ex;
    u32 N = c->N;
    int nsets = c->nsets;
    u32 *sa = P<u32>(c->sa), *lcp = P<u32>(c->t5), *R = P<u32>(c->t1), *dv = P<u32>(c->t4), *prevcl = P<u32>(c->t0);
    int mbits = bits_for((u64)c->mmax - 1);
    TRY(dev_zero(ex, c->set_collected.p, sizeof(u32) * nsets));
    TRY(dev_zero(ex, c->set_suffixfree.p, sizeof(u32) * nsets));
    { WinDepthArgs a{v, sa, lcp, R, dv}; launch_windepth(ex, N, a); }
    { ClKeyArgs a{v, sa, P<u64>(c->keysA), P<u32>(c->valsA), mbits}; launch_clkey(ex, N, a); }
    TRY(sort_pairs(c, N, 0, mbits + CSA_LETTER_BITS));
    { PrevClArgs a{v, sa, P<u64>(c->keysA), P<u32>(c->valsA), prevcl, mbits}; launch_prevcl(ex, N, a); }
    { PlateauArgs a{v, sa, lcp, R, dv, prevcl, P<u32>(c->set_collected), P<u32>(c->set_suffixfree)}; launch_plateau(ex, N, a); }
    c->h_set_collected.assign(nsets, 0); c->h_set_suffixfree.assign(nsets, 0);
    TRY(d2h(ex, c->h_set_collected.data(), c->set_collected.p, sizeof(u32) * nsets));
    TRY(d2h(ex, c->h_set_suffixfree.data(), c->set_suffixfree.p, sizeof(u32) * nsets));
    {   // marked sets: the counts of the literal list walk (rare.cuh)
        std::vector<u32> fl(nsets), rc(nsets), rs(nsets);
        TRY(d2h(ex, fl.data(), c->set_flags.p, sizeof(u32) * nsets));
        bool any = false;
        for (u32 f : fl) any |= (f & CSA_FLAG_RARE) != 0;
        if (any) {
            TRY(d2h(ex, rc.data(), c->rare_collected.p, sizeof(u32) * nsets));
            TRY(d2h(ex, rs.data(), c->rare_suffixfree.p, sizeof(u32) * nsets));
            for (int s = 0; s < nsets; s++)
                if (fl[s] & CSA_FLAG_RARE) { c->h_set_collected[s] = rc[s]; c->h_set_suffixfree[s] = rs[s]; }
        }
    }
    c->have_stats = true;
    return 0;
}

// sequence 0 of every set: its rotations in SA order (sa0), their LCPs (lcp0) and two pyramids of block minima over them --
// all the block order, the literal list walk (rare.cuh) and the block letters need of the reference's tree (Seq0Q)
static int build_pyramid(csa_gpu_ctx *c, DevMem &mem, const u32 *level0, u32 n, Pyramid &py) {
    Exec &ex = c->ex;
    size_t tot = 0;
    for (u32 sz = n; sz > 32;) { sz = (sz + 31) / 32; tot += sz; }
    TRY(dev_alloc(mem, sizeof(u32) * (tot + 32)));
    py.nlev = 1; py.lev[0] = level0; py.size[0] = n;
    u32 *next = P<u32>(mem);
    while (py.size[py.nlev - 1] > 32 && py.nlev < PYR_MAX) {
        u32 nin = py.size[py.nlev - 1], nout = (nin + 31) / 32;
        PyrArgs a{py.lev[py.nlev - 1], next, nin};
        launch_pyr(ex, nout, a);
        py.lev[py.nlev] = next; py.size[py.nlev] = nout; py.nlev++;
        next += nout;
    }
    for (int l = py.nlev; l < PYR_MAX; l++) { py.lev[l] = nullptr; py.size[l] = 0; }
    return 0;
}

static int stage_seq0(csa_gpu_ctx *c, const BatchView &v) {
    Exec &ex = c->ex;
    u32 N0 = c->N0;
    u32 *sa = P<u32>(c->sa), *lcp = P<u32>(c->t5);
    size_t n1 = sizeof(u32) * (size_t)N0;
    DevMem *one[] = {&c->sa0, &c->saidx0, &c->leaf_set, &c->lcp0};
    for (DevMem *m : one) TRY(dev_alloc(*m, n1));
    if (c->use_cover) { // (the colour sort left the SA places of sequence 0 first)
        Seq0TakeArgs a{v, sa, P<u32>(c->saidx0), P<u32>(c->sa0), P<u32>(c->saidx0), P<u32>(c->leaf_set)}; launch_seq0take(ex, N0, a);
    } else {            // the rotations of sequence 0 (a range check) picked out of the suffix array, in its order
#ifdef CSA_EMU
        u32 *flag = P<u32>(c->t4), *idx = P<u32>(c->t1);
        { Seq0FlagArgs a{v, sa, flag, 0u}; launch_seq0flag(ex, c->N, a); }
        TRY((scan_u32<ScanSum, false>(ex, c->ps, flag, idx, c->N)));
        { Seq0EmitArgs a{v, sa, flag, idx, P<u32>(c->sa0), P<u32>(c->saidx0), P<u32>(c->leaf_set), 0u, nullptr, c->N}; launch_seq0emit(ex, c->N, a); }
#else
        const long long nt = ((long long)c->N + CS_TILE - 1) / CS_TILE;
        TRY(dev_alloc(c->ps.chain, sizeof(unsigned long long) * (size_t)(nt + 1)));
        CUDA_TRY(cudaMemsetAsync(c->ps.chain.p, 0, sizeof(unsigned long long) * (size_t)(nt + 1), ex.stream));
        Seq0CompactArgs a{v, sa, P<u32>(c->sa0), P<u32>(c->saidx0), P<u32>(c->leaf_set), (unsigned long long *)c->ps.chain.p, 0u, nullptr};
        PROF_BEGIN(ex, "k_seq0compact", 4.0 * c->N + 12.0 * N0);
        k_seq0compact<<<(unsigned)nt, CS_THREADS, 0, ex.stream>>>((long long)c->N, a);
        PROF_END(ex);
        ex.launches++;
#endif
    }
    { Lcp0Args a{lcp, P<u32>(c->saidx0), P<u32>(c->leaf_set), P<u32>(c->z0), P<u32>(c->lcp0)}; launch_lcp0(ex, N0, a); }
    c->q0.N0 = N0; c->q0.z0 = P<u32>(c->z0); c->q0.leaf_set = P<u32>(c->leaf_set); c->q0.saidx0 = P<u32>(c->saidx0);
    TRY(build_pyramid(c, c->pyr, P<u32>(c->lcp0), N0, c->q0.lcp));
    TRY(build_pyramid(c, c->pyr2, P<u32>(c->sa0), N0, c->q0.pos));
    return 0;
}

static int stage_block_order(csa_gpu_ctx *c, const BatchView &v) {
    Exec &ex = c->ex;
    u32 B = c->B;
    u32 *sa = P<u32>(c->sa);
    // (set, depth descending), stable; then the blocks of equal depth of a set by the DFS that met them
    TRY(dev_alloc(c->blk_leaf, sizeof(u32) * (size_t)B));
    TRY(dev_alloc(c->order, sizeof(u32) * (size_t)B));
    { BlockKeyArgs k{v, sa, c->q0, P<u32>(c->blk_lb), P<u32>(c->blk_depth), P<u32>(c->blk_set), P<u64>(c->keysA), P<u32>(c->valsA), P<u32>(c->blk_leaf)};
      launch_blockkey(ex, B, k); }
    TRY(sort_pairs(c, B, 0, 32 + bits_for((u64)c->nsets - 1)));
    u32 *cstart = P<u32>(c->valsB); // (free between sorts)
    { BlockClassArgs k{P<u64>(c->keysA), cstart}; launch_blockclass(ex, B, k); }
    TRY((scan_u32<ScanMax, true>(ex, c->ps, cstart, cstart, B)));
    TRY(dev_alloc(c->blk_tab, sizeof(u32) * (size_t)B * BR_DEPTHS));
    { BlockTabArgs t{c->q0, P<u32>(c->blk_leaf), P<u32>(c->blk_tab)}; launch_blocktab(ex, (long long)B * BR_DEPTHS, t); }
    { BlockRankArgs r{c->q0, P<u64>(c->keysA), P<u32>(c->valsA), P<u32>(c->blk_leaf), B, cstart, P<u32>(c->blk_tab), P<u32>(c->order)}; launch_blockrank(ex, B, r); }
    return 0;
}

static int stage_chain(csa_gpu_ctx *c, const BatchView &v, int max_interval) {
    Exec &ex = c->ex;
    u32 B = c->B, E = c->E;
    int nsets = c->nsets;
    size_t nb = sizeof(u32) * (size_t)B, ne = sizeof(u32) * (size_t)E;
    DevMem *perblock[] = {&c->o_depth, &c->o_set, &c->succ_lo, &c->succ_hi, &c->next, &c->gap, &c->size, &c->total,
                          &c->interval, &c->inv, &c->f_depth, &c->f_size, &c->f_total, &c->f_interval, &c->f_next, &c->f_leaf, &c->f_set};
    for (DevMem *m : perblock) TRY(dev_alloc(*m, nb));
    DevMem *perelem[] = {&c->o_pos, &c->elem_blk, &c->seghead, &c->f_pos};
    for (DevMem *m : perelem) TRY(dev_alloc(*m, ne));
    u32 *sa = P<u32>(c->sa);
    u32 *set_blk0 = P<u32>(c->set_blk0), *set_pos0 = P<u32>(c->set_pos0);
    { BlockGatherArgs a{v, sa, P<u32>(c->order), P<u32>(c->blk_lb), P<u32>(c->blk_depth), P<u32>(c->blk_set), set_blk0, set_pos0,
                        P<u32>(c->o_depth), P<u32>(c->o_set), P<int>(c->o_pos)};
      launch_blockgather(ex, B, a); }
    { ElemBlkArgs a{v, P<u32>(c->o_set), set_blk0, set_pos0, P<u32>(c->elem_blk)}; launch_elemblk(ex, B, a); }
    int ebits = bits_for(2ull * c->nmax);
    { EndKeyArgs a{v, P<u32>(c->o_depth), P<u32>(c->o_set), P<int>(c->o_pos), set_blk0, set_pos0, P<u32>(c->elem_blk),
                   P<u64>(c->keysA), P<u32>(c->valsA), ebits};
      launch_endkey(ex, E, a); }
    TRY(sort_pairs(c, E, 0, ebits + bits_for((u64)c->M - 1)));
    { SegHeadArgs a{P<u64>(c->keysA), P<u32>(c->seghead), ebits}; launch_seghead(ex, E, a); }
    TRY((scan_u32<ScanMax, true>(ex, c->ps, P<u32>(c->seghead), P<u32>(c->seghead), E)));
    TRY(dev_fill_ff(ex, c->succ_lo.p, nb));
    TRY(dev_zero(ex, c->succ_hi.p, nb));
    { LinkArgs a{v, P<u64>(c->keysA), P<u32>(c->valsA), P<u32>(c->seghead), P<u32>(c->o_depth), ebits, P<u32>(c->succ_lo), P<u32>(c->succ_hi)};
      launch_link(ex, E, a); }
    { RareWalkArgs a{v, sa, P<u32>(c->t5), P<u32>(c->t1), P<u32>(c->set_flags), P<u32>(c->set_neff), P<u32>(c->seq_per), P<u32>(c->t2),
                     set_blk0, set_pos0, P<u32>(c->o_depth), P<int>(c->o_pos),
                     {P<u32>(c->keysA), P<u32>(c->keysA) + c->N, P<u32>(c->keysB), P<u32>(c->keysB) + c->N, P<u32>(c->valsA), P<u32>(c->valsB)},
                     P<u32>(c->succ_lo), P<u32>(c->succ_hi)};
      launch_rarewalk(ex, nsets, a); }
    { GapArgs a{v, P<u32>(c->succ_lo), P<u32>(c->succ_hi), P<u32>(c->o_depth), P<u32>(c->o_set), P<int>(c->o_pos), set_blk0, set_pos0,
                max_interval, P<int>(c->next), P<int>(c->gap)};
      launch_gap(ex, B, a); }
    TRY(dev_zero(ex, c->size.p, nb)); TRY(dev_zero(ex, c->total.p, nb)); TRY(dev_zero(ex, c->interval.p, nb));
    {
        ChainArgs a{set_blk0, P<u32>(c->o_depth), P<int>(c->next), P<int>(c->gap), P<int>(c->size), P<int>(c->total), P<int>(c->interval),
                    P<u32>(c->set_nchains), P<u32>(c->set_flags), 0};
#ifndef CSA_EMU
        // sets with more blocks than k_chain holds in shared memory: k_chain_big (walk order in shared memory, sums in parallel)
        std::vector<u32> big, evb;
        u64 nev = 0;
        u32 maxb = 0;
        for (int s = 0; s < nsets; s++) {
            const u32 nbk = c->h_set_nblocks[s];
            if (nbk > CH_CAP && nbk <= CHB_MAX) { big.push_back((u32)s); evb.push_back((u32)nev); nev += 4ull * nbk; maxb = std::max(maxb, nbk); }
        }
        if (!big.empty() && nev < (1ull << 32) && !c->no_chain_big) {
            a.skip_big = 1;
            TRY(dev_alloc(c->chb_sets, sizeof(u32) * big.size())); TRY(dev_alloc(c->chb_evbase, sizeof(u32) * big.size()));
            TRY(dev_alloc(c->chb_events, sizeof(u64) * (size_t)nev)); TRY(dev_alloc(c->chb_work, 9 * nb));
            TRY(dev_alloc(c->chb_redo, sizeof(u32) * (nsets + 1)));
            TRY(h2d(ex, c->chb_sets.p, big.data(), sizeof(u32) * big.size()));
            TRY(h2d(ex, c->chb_evbase.p, evb.data(), sizeof(u32) * big.size()));
            TRY(dev_zero(ex, c->chb_redo.p, sizeof(u32) * (nsets + 1)));
            ChainBigArgs g{a, P<u32>(c->chb_sets), P<unsigned long long>(c->chb_events), P<u32>(c->chb_evbase),
                           P<int>(c->chb_work), P<int>(c->chb_work) + B, P<int>(c->chb_work) + 2 * (size_t)B, P<u32>(c->chb_work) + 3 * (size_t)B,
                           P<u32>(c->chb_work) + 4 * (size_t)B, P<u32>(c->chb_work) + 5 * (size_t)B, P<u32>(c->chb_work) + 6 * (size_t)B,
                           P<u32>(c->chb_work) + 7 * (size_t)B, P<u32>(c->chb_work) + 8 * (size_t)B,
                           P<u32>(c->chb_redo)};
            TRY(launch_chain_big(ex, (u32)big.size(), 3 * (size_t)maxb + 16, g));
        }
#endif
        launch_chain(ex, nsets, a);
#ifndef CSA_EMU
        if (a.skip_big) { // a set whose sums contradicted the walk order (see k_chain_big): the literal walk redoes it
            std::vector<u32> redo(nsets);
            TRY(d2h(ex, redo.data(), c->chb_redo.p, sizeof(u32) * nsets));
            bool any = false;
            for (u32 r : redo) any |= r != 0;
            c->chain_redone = any ? 1 : 0;
            if (any) {
                TRY(dev_zero(ex, c->size.p, nb)); TRY(dev_zero(ex, c->total.p, nb)); TRY(dev_zero(ex, c->interval.p, nb));
                a.skip_big = 0;
                launch_chain(ex, nsets, a);
            }
        }
#endif
    }
    { SizeKeyArgs a{P<u32>(c->o_set), P<int>(c->size), P<u64>(c->keysA), P<u32>(c->valsA)}; launch_sizekey(ex, B, a); }
    TRY(sort_pairs(c, B, 0, 32 + bits_for((u64)c->nsets - 1)));
    { InvArgs a{P<u32>(c->valsA), P<u32>(c->inv)}; launch_inv(ex, B, a); }
    { FinalArgs a{v, P<u32>(c->valsA), P<u32>(c->inv), P<u32>(c->o_depth), P<u32>(c->o_set), P<int>(c->o_pos), P<int>(c->size),
                  P<int>(c->total), P<int>(c->interval), P<int>(c->next), set_blk0, set_pos0, P<int>(c->f_depth), P<int>(c->f_size),
                  P<int>(c->f_total), P<int>(c->f_interval), P<int>(c->f_next), P<int>(c->f_pos),
                  P<u32>(c->order), P<u32>(c->blk_leaf), P<u32>(c->f_leaf)};
      launch_final(ex, B, a); }
    TRY(d2d(ex, c->f_set.p, c->o_set.p, nb)); // (sets are contiguous in both orders: block i of the final list belongs to o_set[i])
    return 0;
}

// phase 0: the whole path; 1: up to this rank's bucket of the suffix array (csa_gpu_shard_begin); 2: the rest
static void fold_profile(csa_gpu_ctx *c) { // the event pairs of this run into per-kernel sums
#ifndef CSA_EMU
    if (!c->ex.prof) return;
    c->prof_sum.clear();
    for (ProfRec &r : c->prof.recs) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        size_t j = 0;
        while (j < c->prof_sum.size() && c->prof_sum[j].name != r.name) j++;
        if (j == c->prof_sum.size()) c->prof_sum.push_back({r.name, 0, 0.0, 0.0});
        c->prof_sum[j].launches++; c->prof_sum[j].ms += ms; c->prof_sum[j].bytes += r.bytes;
    }
    c->prof.recs.clear();
    c->prof.pool_used = 0;
#else
    (void)c;
#endif
}

static int run_phases(csa_gpu_ctx *c, int max_interval, unsigned flags, int phase) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (!c->uploaded) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_batch_run before csa_gpu_batch_upload");
#ifndef CSA_EMU
    CUDA_TRY(cudaSetDevice(c->device));
#endif
    Exec &ex = c->ex;
    if (phase != 2) ex.launches = 0;
    c->ran = false;
    u32 N = c->N;
    int nsets = c->nsets;
    size_t n4 = sizeof(u32) * (size_t)N;
    // (the sort buffers also serve the sequence-0 tree: 2 N0 nodes, more than N when a set's first sequence is its longest by far)
    size_t nk = sizeof(u32) * (size_t)std::max<u64>(N, 2ull * c->N0);
    TRY(dev_alloc(c->keysA, 2 * nk)); TRY(dev_alloc(c->keysB, 2 * nk));
    TRY(dev_alloc(c->valsA, nk)); TRY(dev_alloc(c->valsB, nk)); TRY(dev_alloc(c->sa, nk)); // (sa and valsA trade places after the suffix sort)
    TRY(dev_alloc(c->t0, n4)); TRY(dev_alloc(c->t1, n4)); TRY(dev_alloc(c->t2, n4)); TRY(dev_alloc(c->t3, n4)); TRY(dev_alloc(c->t4, n4)); TRY(dev_alloc(c->t5, n4));
    TRY(dev_alloc(c->counter, 256));
    DevMem *perset[] = {&c->set_nblocks, &c->set_blk0, &c->set_pos0, &c->set_flags, &c->set_nchains, &c->set_cyclic, &c->firstmax,
                        &c->set_collected, &c->set_suffixfree, &c->set_neff, &c->rare_collected, &c->rare_suffixfree};
    TRY(dev_alloc(c->seq_per, sizeof(u32) * (size_t)c->M));
    for (DevMem *m : perset) TRY(dev_alloc(*m, sizeof(u32) * (nsets + 1)));
    TRY(dev_alloc(c->rotations, sizeof(int) * (size_t)c->M));
    TRY(dev_alloc(c->shard_bounds, sizeof(u32) * ((size_t)c->shard_nranks + 2)));
    BatchView v = view_of(c);

    u32 *any_other = P<u32>(c->counter) + 8;
    if (phase != 2) {
        mark(c, 0);
        TRY(dev_zero(ex, any_other, sizeof(u32)));
        { EncodeArgs a{v, P<unsigned char>(c->raw), any_other}; launch_encode(ex, N, a); }
        { PackArgs a{v}; launch_pack(ex, (long long)c->TW, a); }
    }
    TRY(stage_suffix_array(c, v, phase));
    if (phase == 1) { c->shard_phase = 1; return exec_sync(ex); }
    c->shard_phase = 0;
    mark(c, 1);
    if (c->lcp_state == 2) { // the word sort gave all but the groups the doubling rounds finished
        LcpFixArgs a{v, P<u32>(c->sa), P<u32>(c->t5), any_other};
        launch_lcpfix(ex, N, a);
    } else if (c->lcp_state == 0) {   // how long are the matches?  a sample of pairs decides between the two LCP kernels
        const u32 nsample = 4096, stride = N / nsample + 1;
        unsigned long long *sum = (unsigned long long *)(P<u32>(c->counter) + 10), hsum = 0;
        TRY(dev_zero(ex, sum, sizeof(*sum)));
        { LcpDirectArgs a{v, P<u32>(c->sa), nullptr, any_other, stride, sum}; launch_lcpdirect(ex, (N + stride - 1) / stride, a); }
        TRY(d2h(ex, &hsum, sum, sizeof(hsum)));
        const double mean = (double)hsum / (double)((N + stride - 1) / stride);
        c->lcp_mean_sample = mean;
        if (mean < 96.0 && !c->force_kasai) {
            LcpDirectArgs a{v, P<u32>(c->sa), P<u32>(c->t5), any_other, 1, nullptr};
            launch_lcpdirect(ex, N, a);
        } else {
            { IsaArgs a{P<u32>(c->sa), P<u32>(c->t1)}; launch_isa(ex, N, a); }
            { LcpArgs a{v, P<u32>(c->sa), P<u32>(c->t1), P<u32>(c->t5), any_other}; launch_lcp(ex, ((long long)N + LCP_CHUNK - 1) / LCP_CHUNK, a); }
        }
    }
    mark(c, 2);
    TRY(dev_zero(ex, c->set_flags.p, sizeof(u32) * nsets));
    // sets whose tree is not their suffix array (rare.cuh): marked here, redone by one thread each further down
    { LeafScanArgs a{v, P<u32>(c->t5), P<u32>(c->set_flags), c->batch_nmin, 0u}; launch_leafscan(ex, N, a); }
    { RareCollapseArgs a{v, P<u32>(c->sa), P<u32>(c->t5), P<u32>(c->set_flags), P<u32>(c->t2), P<u32>(c->set_neff), P<u32>(c->seq_per)};
      launch_rarecollapse(ex, nsets, a); }
    c->use_cover = (flags & CSA_GPU_FLAG_STATS) || c->mmax > BF2_MAXM || c->force_cover;
    TRY(stage_common_blocks(c, v));
    TRY(stage_seq0(c, v));
    { RareBlocksArgs a{v, P<u32>(c->sa), P<u32>(c->t5), P<u32>(c->t1), c->use_cover ? 1 : 0, P<u32>(c->set_flags), P<u32>(c->set_neff), P<u32>(c->seq_per), P<u32>(c->t2),
                       c->q0,
                       {P<u32>(c->keysA), P<u32>(c->keysA) + N, P<u32>(c->keysB), P<u32>(c->keysB) + N, P<u32>(c->valsA), P<u32>(c->valsB)},
                       P<u32>(c->t0), P<u32>(c->t3), P<u32>(c->rare_collected), P<u32>(c->rare_suffixfree)};
      launch_rareblocks(ex, nsets, a); }
    TRY(stage_emit_blocks(c, v));
    c->have_stats = false;
    if (flags & CSA_GPU_FLAG_STATS) TRY(stage_stats(c, v));
    mark(c, 3);
    TRY(stage_block_order(c, v));
    mark(c, 4);
    TRY(stage_chain(c, v, max_interval));
    { RotArgs a{v, P<u32>(c->set_blk0), P<u32>(c->set_pos0), P<int>(c->f_pos), P<int>(c->f_next), P<int>(c->rotations), P<u32>(c->set_cyclic)};
      launch_rot(ex, nsets, a); }
    mark(c, 5);
    TRY(exec_sync(ex));
#ifndef CSA_EMU
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) CSA_FAIL(CSA_GPU_ECUDA, "kernel failure: %s", cudaGetErrorString(e));
        for (int i = 0; i < 5; i++) cudaEventElapsedTime(&c->tm.ms[i], c->tm.ev[i], c->tm.ev[i + 1]);
        cudaEventElapsedTime(&c->tm.ms[5], c->tm.ev[0], c->tm.ev[5]);
    }
#endif
    c->launches = ex.launches;
    c->ran = true;
    fold_profile(c);
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_batch_run(csa_gpu_ctx *c, int max_interval, unsigned flags) {
    if (c) { c->shard_rank = 0; c->shard_nranks = 1; }
    return run_phases(c, max_interval, flags, 0);
}

// ---- one batch, the suffix-array stage sharded over the ranks of a job ----------------------------------------
extern "C" int csa_gpu_shard_begin(csa_gpu_ctx *c, int rank, int nranks) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (nranks < 1 || rank < 0 || rank >= nranks) CSA_FAIL(CSA_GPU_EINVAL, "bad rank %d of %d", rank, nranks);
    c->shard_rank = rank; c->shard_nranks = nranks; c->shard_sa_swapped = false;
    return run_phases(c, 0, 0, 1);
}

extern "C" int csa_gpu_shard_advice(csa_gpu_ctx *c, int nranks) {
    if (!c || !c->uploaded) return 1;
    const bool carried = c->nsets == 1 && c->carry_mode != 2 && c->max_set_bases > WS_LARGE_SET && !c->shard_full_sort;
    return (!carried || nranks >= CSA_GPU_SHARD_MIN_RANKS) ? 1 : 0;
}

extern "C" int csa_gpu_shard_view(csa_gpu_ctx *c, csa_gpu_shard_info *out) {
    if (!c || !out) CSA_FAIL(CSA_GPU_EINVAL, "null argument");
    if (c->shard_phase != 1) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_shard_view before csa_gpu_shard_begin");
    out->sa = c->valsA.p; out->head = c->t0.p; out->lcp = c->t5.p; out->left = c->t4.p;
    out->n = c->N;
    out->bounds = c->h_shard_bounds.data();
    out->nleft = c->ws_left[0]; out->left_suffixes = c->ws_left[1]; out->min_depth = c->ws_left[2]; out->max_group = c->ws_left[3];
    out->own_sort = c->shard_own_sort ? 1u : 0u;
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_shard_finish(csa_gpu_ctx *c, int max_interval, unsigned flags, unsigned nleft, unsigned left_suffixes,
                                    unsigned min_depth, unsigned max_group) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (c->shard_phase != 1) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_shard_finish before csa_gpu_shard_begin");
    c->ws_left[0] = nleft; c->ws_left[1] = left_suffixes; c->ws_left[2] = min_depth; c->ws_left[3] = max_group;
    if (c->shard_sa_swapped) { std::swap(c->sa, c->valsA); c->shard_sa_swapped = false; } // (csa_gpu_shard_blocks_begin ran, then the caller fell back)
    return run_phases(c, max_interval, flags, 2);
}

// ---- ... and the block stages of a rank's own range (see include/csa_gpu.h) -------------------------------------------------
extern "C" int csa_gpu_shard_blocks_begin(csa_gpu_ctx *c, unsigned flags, csa_gpu_shard_blocks *out) {
    if (!c || !out) CSA_FAIL(CSA_GPU_EINVAL, "null argument");
    if (c->shard_phase != 1) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_shard_blocks_begin before csa_gpu_shard_begin");
    if (!c->shard_own_sort || c->nsets != 1 || c->mmax > 64 || (flags & CSA_GPU_FLAG_STATS) || c->ws_left[0])
        CSA_FAIL(CSA_GPU_ESTATE, "the block stages shard for ONE set of up to 64 sequences whose buckets were sorted by their ranks, "
                                 "without the counts of csamsa.c:332,338: take csa_gpu_shard_finish (own sort %d, sets %d, "
                                 "sequences %d, flags %u, groups left %u)", (int)c->shard_own_sort, c->nsets, c->mmax, flags, c->ws_left[0]);
#ifndef CSA_EMU
    CUDA_TRY(cudaSetDevice(c->device));
#endif
    Exec &ex = c->ex;
    BatchView v = view_of(c);
    const u32 N = c->N, m = c->mmax, lo = c->h_shard_bounds[c->shard_rank], hi = c->h_shard_bounds[c->shard_rank + 1], n = hi - lo;
    memset(out, 0, sizeof(*out));
    out->m = m; out->head_min = out->tail_min = 0xFFFFFFFFu;
    if (!c->shard_sa_swapped) { std::swap(c->sa, c->valsA); c->shard_sa_swapped = true; } // (as the end of the suffix-array stage does)
    u32 *sa = P<u32>(c->sa), *lcp = P<u32>(c->t5), *isblock = P<u32>(c->t0), *depth = P<u32>(c->t3), *bidx = P<u32>(c->t4);
    u32 *any_other = P<u32>(c->counter) + 8, *cnt = P<u32>(c->counter) + 14;
    if (n == 0) return CSA_GPU_OK;
    // the LCPs at the two borders of the range: no rank has seen both keys
    if (lo > 0) { LcpAtArgs a{{v, sa, lcp, any_other, 1, nullptr}, lo}; launch_lcpat(ex, 1, a); }
    if (hi < N) { LcpAtArgs a{{v, sa, lcp, any_other, 1, nullptr}, hi}; launch_lcpat(ex, 1, a); }
    TRY(dev_zero(ex, c->set_flags.p, sizeof(u32) * c->nsets));
    { LeafScanArgs a{v, lcp, P<u32>(c->set_flags), c->batch_nmin, lo}; launch_leafscan(ex, n, a); }
    { BlockFind2Args a{v, sa, lcp, isblock, depth, m, lo}; launch_blockfind2(ex, n, a); }
    TRY((scan_u32<ScanSum, false>(ex, c->ps, isblock + lo, bidx + lo, n)));
    // the rotations of sequence 0 in the range, in suffix-array order
    const u32 cap0 = std::min(n, c->N0);
    TRY(dev_alloc(c->sh_sa0, sizeof(u32) * (size_t)cap0)); TRY(dev_alloc(c->sh_saidx0, sizeof(u32) * (size_t)cap0));
    TRY(dev_alloc(c->sh_lcp0, sizeof(u32) * (size_t)cap0));
    u32 *loc_set = P<u32>(c->t1); // (leaf_set of the range's rotations: all of set 0; scratch)
    TRY(dev_zero(ex, cnt, 4 * sizeof(u32)));
#ifdef CSA_EMU
    {
        u32 *flag = P<u32>(c->keysA), *idx = P<u32>(c->keysA) + N;
        { Seq0FlagArgs a{v, sa, flag, lo}; launch_seq0flag(ex, n, a); }
        TRY((scan_u32<ScanSum, false>(ex, c->ps, flag, idx, n)));
        { Seq0EmitArgs a{v, sa, flag, idx, P<u32>(c->sh_sa0), P<u32>(c->sh_saidx0), loc_set, lo, cnt, n}; launch_seq0emit(ex, n, a); }
    }
#else
    {
        const long long nt = ((long long)n + CS_TILE - 1) / CS_TILE;
        TRY(dev_alloc(c->ps.chain, sizeof(unsigned long long) * (size_t)(nt + 1)));
        CUDA_TRY(cudaMemsetAsync(c->ps.chain.p, 0, sizeof(unsigned long long) * (size_t)(nt + 1), ex.stream));
        Seq0CompactArgs a{v, sa, P<u32>(c->sh_sa0), P<u32>(c->sh_saidx0), loc_set, (unsigned long long *)c->ps.chain.p, lo, cnt};
        PROF_BEGIN(ex, "k_seq0compact", 4.0 * n);
        k_seq0compact<<<(unsigned)nt, CS_THREADS, 0, ex.stream>>>((long long)n, a);
        PROF_END(ex);
        ex.launches++;
    }
#endif
    u32 hv[4] = {0, 0, 0, 0}; // [0] rotations of sequence 0, [1] running count at the last place, [2] is the last place a block, [3] flags
    TRY(d2d(ex, cnt + 1, bidx + (hi - 1), sizeof(u32))); TRY(d2d(ex, cnt + 2, isblock + (hi - 1), sizeof(u32)));
    TRY(d2d(ex, cnt + 3, c->set_flags.p, sizeof(u32)));
    TRY(d2h(ex, hv, cnt, sizeof(hv)));
    const u32 n0 = hv[0], nblk = hv[1] + hv[2];
    out->n0 = n0; out->nblk = nblk; out->rare = (hv[3] & CSA_FLAG_RARE) ? 1u : 0u;
    TRY(dev_alloc(c->sh_rec, sizeof(u32) * (size_t)std::max(nblk, 1u) * (2 + m)));
    { BlkRecArgs a{sa, isblock, bidx + 0, depth, lo, m, P<u32>(c->sh_rec)}; launch_blkrecpack(ex, n, a); }
    // their LCPs; what the first one lacks (the places before the range) and what lies behind the last one: two minima
    u32 *mins = P<u32>(c->counter) + 18;
    TRY(dev_fill_ff(ex, mins, 2 * sizeof(u32)));
    if (n0) {
        u32 ends[2] = {0, 0};
        TRY(d2h(ex, ends, c->sh_saidx0.p, sizeof(u32)));
        TRY(d2h(ex, ends + 1, P<u32>(c->sh_saidx0) + (n0 - 1), sizeof(u32)));
        std::vector<u32> zz = {0u, n0};
        TRY(h2d(ex, P<u32>(c->counter) + 30, zz.data(), 2 * sizeof(u32))); // (one "set" of n0 leaves for k_lcp0)
        TRY(exec_sync(ex));
        { Lcp0Args a{lcp, P<u32>(c->sh_saidx0), loc_set, P<u32>(c->counter) + 30, P<u32>(c->sh_lcp0)}; launch_lcp0(ex, n0, a); }
        { RangeMinArgs a{lcp, lo, mins}; launch_rangemin(ex, (long long)ends[0] - lo + 1, a); }
        if (ends[1] + 1 < hi) { RangeMinArgs a{lcp, ends[1] + 1, mins + 1}; launch_rangemin(ex, (long long)hi - ends[1] - 1, a); }
    } else {
        RangeMinArgs a{lcp, lo, mins}; launch_rangemin(ex, n, a);
    }
    u32 hm[2];
    TRY(d2h(ex, hm, mins, sizeof(hm)));
    out->head_min = hm[0]; out->tail_min = hm[1];
    out->blkrec = c->sh_rec.p; out->sa0 = c->sh_sa0.p; out->saidx0 = c->sh_saidx0.p; out->lcp0 = c->sh_lcp0.p;
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_shard_blocks_buffers(csa_gpu_ctx *c, unsigned total_blocks, unsigned total_n0, void **blkrec, void **sa0,
                                            void **saidx0, void **lcp0) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (c->shard_phase != 1) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_shard_blocks_buffers before csa_gpu_shard_begin");
    if (total_n0 != c->N0) CSA_FAIL(CSA_GPU_EINVAL, "the ranks hold %u rotations of sequence 0, the set has %u", total_n0, c->N0);
#ifndef CSA_EMU
    CUDA_TRY(cudaSetDevice(c->device));
#endif
    TRY(dev_alloc(c->sh_allrec, sizeof(u32) * (size_t)std::max(total_blocks, 1u) * (2 + c->mmax)));
    size_t n1 = sizeof(u32) * (size_t)c->N0;
    DevMem *one[] = {&c->sa0, &c->saidx0, &c->leaf_set, &c->lcp0};
    for (DevMem *m : one) TRY(dev_alloc(*m, n1));
    if (blkrec) *blkrec = c->sh_allrec.p;
    if (sa0) *sa0 = c->sa0.p;
    if (saidx0) *saidx0 = c->saidx0.p;
    if (lcp0) *lcp0 = c->lcp0.p;
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_shard_blocks_finish(csa_gpu_ctx *c, int max_interval, unsigned flags, unsigned total_blocks, unsigned total_n0) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (c->shard_phase != 1 || !c->shard_sa_swapped) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_shard_blocks_finish before csa_gpu_shard_blocks_begin");
    if (total_n0 != c->N0 || (flags & CSA_GPU_FLAG_STATS)) CSA_FAIL(CSA_GPU_EINVAL, "bad argument");
#ifndef CSA_EMU
    CUDA_TRY(cudaSetDevice(c->device));
#endif
    Exec &ex = c->ex;
    BatchView v = view_of(c);
    const u32 B = total_blocks, m = c->mmax;
    mark(c, 1); mark(c, 2);
    c->B = B; c->E = B * m;
    c->use_cover = false; c->have_stats = false;
    TRY(dev_alloc(c->blk_lb, sizeof(u32) * (size_t)B)); TRY(dev_alloc(c->blk_depth, sizeof(u32) * (size_t)B));
    TRY(dev_alloc(c->blk_set, sizeof(u32) * (size_t)B));
    { BlkUnpackArgs a{P<u32>(c->sh_allrec), m, P<u32>(c->sa), P<u32>(c->blk_lb), P<u32>(c->blk_depth), P<u32>(c->blk_set)};
      launch_blkrecunpack(ex, B, a); }
    c->h_set_nblocks.assign(1, B);
    c->h_set_blk0 = {0u, B}; c->h_set_pos0 = {0u, c->E};
    TRY(h2d(ex, c->set_nblocks.p, c->h_set_nblocks.data(), sizeof(u32)));
    TRY(h2d(ex, c->set_blk0.p, c->h_set_blk0.data(), 2 * sizeof(u32)));
    TRY(h2d(ex, c->set_pos0.p, c->h_set_pos0.data(), 2 * sizeof(u32)));
    TRY(dev_zero(ex, c->set_flags.p, sizeof(u32) * c->nsets));
    TRY(dev_zero(ex, c->leaf_set.p, sizeof(u32) * (size_t)c->N0));
    c->q0.N0 = c->N0; c->q0.z0 = P<u32>(c->z0); c->q0.leaf_set = P<u32>(c->leaf_set); c->q0.saidx0 = P<u32>(c->saidx0);
    TRY(build_pyramid(c, c->pyr, P<u32>(c->lcp0), c->N0, c->q0.lcp));
    TRY(build_pyramid(c, c->pyr2, P<u32>(c->sa0), c->N0, c->q0.pos));
    mark(c, 3);
    TRY(stage_block_order(c, v));
    mark(c, 4);
    TRY(stage_chain(c, v, max_interval));
    { RotArgs a{v, P<u32>(c->set_blk0), P<u32>(c->set_pos0), P<int>(c->f_pos), P<int>(c->f_next), P<int>(c->rotations), P<u32>(c->set_cyclic)};
      launch_rot(ex, c->nsets, a); }
    mark(c, 5);
    TRY(exec_sync(ex));
#ifndef CSA_EMU
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) CSA_FAIL(CSA_GPU_ECUDA, "kernel failure: %s", cudaGetErrorString(e));
        for (int i = 0; i < 5; i++) cudaEventElapsedTime(&c->tm.ms[i], c->tm.ev[i], c->tm.ev[i + 1]);
        cudaEventElapsedTime(&c->tm.ms[5], c->tm.ev[0], c->tm.ev[5]);
    }
#endif
    c->launches = ex.launches;
    c->shard_phase = 0;
    c->ran = true;
    fold_profile(c);
    return CSA_GPU_OK;
}

// ---- the same from ONE process that drives several GPUs (the C host: no NCCL, no Python) -------------------------
// One host thread per GPU for the two compute phases; the bucket exchange by peer copies (NVLink when the GPUs
// see each other), the lists of left-over groups through the host (they are tiny).
struct csa_gpu_multi {
    std::vector<csa_gpu_ctx *> ctx;
    std::string err;
};

extern "C" int csa_gpu_multi_create(int ngpus, const int *devices, csa_gpu_multi **out) {
    if (!out || ngpus < 1) CSA_FAIL(CSA_GPU_EINVAL, "csa_gpu_multi_create: bad argument");
    *out = nullptr;
    csa_gpu_multi *m = new (std::nothrow) csa_gpu_multi();
    if (!m) CSA_FAIL(CSA_GPU_ENOMEM, "out of host memory");
    for (int i = 0; i < ngpus; i++) {
        csa_gpu_ctx *c = nullptr;
        int rc = csa_gpu_create(devices ? devices[i] : i, &c);
        if (rc) { for (csa_gpu_ctx *x : m->ctx) csa_gpu_destroy(x); delete m; return rc; }
        m->ctx.push_back(c);
    }
#ifndef CSA_EMU
    for (int i = 0; i < ngpus; i++) // direct loads/stores between the GPUs where the box allows them
        for (int j = 0; j < ngpus; j++) {
            if (i == j || m->ctx[i]->device == m->ctx[j]->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->ctx[i]->device, m->ctx[j]->device);
            if (can) { cudaSetDevice(m->ctx[i]->device); cudaDeviceEnablePeerAccess(m->ctx[j]->device, 0); cudaGetLastError(); }
        }
#endif
    *out = m;
    return CSA_GPU_OK;
}

extern "C" void csa_gpu_multi_destroy(csa_gpu_multi *m) {
    if (!m) return;
    for (csa_gpu_ctx *c : m->ctx) csa_gpu_destroy(c);
    delete m;
}

extern "C" int csa_gpu_multi_size(csa_gpu_multi *m) { return m ? (int)m->ctx.size() : 0; }
extern "C" csa_gpu_ctx *csa_gpu_multi_ctx(csa_gpu_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

// run f(rank) on one host thread per context; the first failure's code and message come back
template <class F> static int multi_parallel(csa_gpu_multi *m, F f) {
    const int R = (int)m->ctx.size();
    std::vector<int> rc(R, 0);
    std::vector<std::string> msg(R);
    std::vector<std::thread> th;
    for (int r = 0; r < R; r++)
        th.emplace_back([&, r]() { rc[r] = f(r); if (rc[r]) msg[r] = g_csa_err; });
    for (std::thread &t : th) t.join();
    for (int r = 0; r < R; r++)
        if (rc[r]) CSA_FAIL(rc[r], "GPU %d of %d: %s", r, R, msg[r].c_str());
    return 0;
}

static int multi_copy(csa_gpu_ctx *dst, void *dp, csa_gpu_ctx *src, const void *sp, size_t bytes) {
    if (!bytes) return 0;
#ifdef CSA_EMU
    (void)dst; (void)src;
    memcpy(dp, sp, bytes);
#else
    CUDA_TRY(cudaMemcpyPeerAsync(dp, dst->device, sp, src->device, bytes, dst->ex.stream));
#endif
    return 0;
}

extern "C" int csa_gpu_multi_batch_rotations(csa_gpu_multi *m, int nsets, const int *set_start, const char *const *texts,
                                             const int *textsizes, int max_interval, unsigned flags, int *rotations,
                                             csa_gpu_set_info *info) {
    if (!m || m->ctx.empty()) CSA_FAIL(CSA_GPU_EINVAL, "null argument");
    const int R = (int)m->ctx.size();
    if (R == 1) return csa_gpu_batch_rotations(m->ctx[0], nsets, set_start, texts, textsizes, max_interval, flags, rotations, info);
    // every GPU: the whole batch, then its own bucket of the suffix array
    TRY(csa_gpu_batch_upload(m->ctx[0], nsets, set_start, texts, textsizes));
    if (!csa_gpu_shard_advice(m->ctx[0], R)) { // one set of whole genomes on few GPUs: one GPU's carried sort is the fastest way
        TRY(csa_gpu_batch_run(m->ctx[0], max_interval, flags));
        return csa_gpu_batch_download(m->ctx[0], rotations, info);
    }
    TRY(multi_parallel(m, [&](int r) {
        int rc = r ? csa_gpu_batch_upload(m->ctx[r], nsets, set_start, texts, textsizes) : 0;
        return rc ? rc : csa_gpu_shard_begin(m->ctx[r], r, R);
    }));
    std::vector<csa_gpu_shard_info> v(R);
    for (int r = 0; r < R; r++) TRY(csa_gpu_shard_view(m->ctx[r], &v[r]));
    const unsigned *b = v[0].bounds;
    for (int r = 1; r < R; r++)
        for (int q = 0; q <= R; q++)
            if (v[r].bounds[q] != b[q]) CSA_FAIL(CSA_GPU_ECUDA, "bucket borders differ between GPUs (%u != %u)", v[r].bounds[q], b[q]);
    // what the bucket sorts left, through the host
    unsigned nleft = 0, left_suffixes = 0, min_depth = 0xFFFFFFFFu, max_group = 0;
    std::vector<unsigned long long> left;
    for (int r = 0; r < R; r++) {
        const size_t at = left.size();
        left.resize(at + v[r].nleft);
        if (v[r].nleft) {
#ifndef CSA_EMU
            CUDA_TRY(cudaSetDevice(m->ctx[r]->device));
#endif
            TRY(d2h(m->ctx[r]->ex, left.data() + at, v[r].left, sizeof(unsigned long long) * v[r].nleft));
        }
        nleft += v[r].nleft; left_suffixes += v[r].left_suffixes;
        min_depth = std::min(min_depth, v[r].min_depth); max_group = std::max(max_group, v[r].max_group);
    }
    // every bucket to every other GPU: suffix array and LCP, the group heads only when groups were left
    for (int q = 0; q < R; q++) {
#ifndef CSA_EMU
        CUDA_TRY(cudaSetDevice(m->ctx[q]->device));
#endif
        for (int r = 0; r < R; r++) {
            if (r == q || b[r + 1] == b[r]) continue;
            const size_t off = (size_t)b[r] * sizeof(u32), bytes = (size_t)(b[r + 1] - b[r]) * sizeof(u32);
            TRY(multi_copy(m->ctx[q], (char *)v[q].sa + off, m->ctx[r], (const char *)v[r].sa + off, bytes));
            TRY(multi_copy(m->ctx[q], (char *)v[q].lcp + off, m->ctx[r], (const char *)v[r].lcp + off, bytes));
            if (nleft) TRY(multi_copy(m->ctx[q], (char *)v[q].head + off, m->ctx[r], (const char *)v[r].head + off, bytes));
        }
        if (nleft) TRY(h2d(m->ctx[q]->ex, v[q].left, left.data(), sizeof(unsigned long long) * left.size()));
    }
    for (int q = 0; q < R; q++) {
#ifndef CSA_EMU
        CUDA_TRY(cudaSetDevice(m->ctx[q]->device));
#endif
        TRY(exec_sync(m->ctx[q]->ex));
    }
    // every GPU: the rest of the path (the results are the same everywhere; they are read from GPU 0)
    TRY(multi_parallel(m, [&](int r) { return csa_gpu_shard_finish(m->ctx[r], max_interval, flags, nleft, left_suffixes, min_depth, max_group); }));
    return csa_gpu_batch_download(m->ctx[0], rotations, info);
}

// tests: force every doubling round down the device-wide radix path (1) or let the tiles decide (0);
// rounds[0], rounds[1] = rounds of the last run that took the tile path / the device-wide path
extern "C" int csa_gpu_debug_rounds(csa_gpu_ctx *c, int force_global, int rounds[2]) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    // force_global: 0 free choice (word sort, then group lists while groups are small), 1 device-wide rounds only,
    // 2 tile rounds with doubling only (and the text-order LCP kernel), 3 tile rounds, quadrupling allowed,
    // 4 word sort stopped after two words (the rest by doubling rounds), 5 free choice without the word sort
    if (force_global >= 0) {
        c->round_mode = force_global;
        c->ws_depth_cap = force_global == 4 ? 80u : WS_DEPTH_CAP;
        c->ws_force = force_global == 6; // 6: word sort whatever the groups look like
        c->no_chain_big = force_global == 7; // 7: free choice, but long block lists walked by one thread (k_chain) as short ones are
        c->shard_full_sort = force_global == 8; // 8: sharded runs of one set sort the whole set on every rank (as batches of sets do)
        c->force_cover = force_global == 9;     // 9: free choice, blocks always through the cover array R[]
        c->carry_mode = force_global == 10 || force_global == 12 ? 1 : force_global == 11 ? 2 : 0; // 10: word sort, carried, whatever the sets look like; 11: free choice, never carried
        c->cy_nopack = force_global == 12; // 12: as 10, with the group table laid out as for batches of 2^27 suffixes and more
        if (force_global == 10 || force_global == 12) c->ws_force = true;
        if (force_global >= 6) c->round_mode = 0;
        c->force_global_rounds = force_global == 1; c->no_quad_rounds = force_global == 2; c->force_kasai = force_global == 2;
    }
    if (rounds) { rounds[0] = c->rounds_tiled + c->rounds_quad + c->rounds_list + c->ws_runs; rounds[1] = c->rounds_global; }
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_profile_enable(csa_gpu_ctx *c, int on) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    c->ex.prof = on ? &c->prof : nullptr;
    if (!on) c->prof_sum.clear();
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_profile_count(csa_gpu_ctx *c) { return c ? (int)c->prof_sum.size() : 0; }

extern "C" int csa_gpu_profile_get(csa_gpu_ctx *c, int i, char *name, int name_cap, long long *launches, double *ms,
                                   double *bytes) {
    if (!c || i < 0 || i >= (int)c->prof_sum.size()) CSA_FAIL(CSA_GPU_EINVAL, "no such profile row");
    if (name && name_cap > 0) snprintf(name, (size_t)name_cap, "%s", c->prof_sum[i].name.c_str());
    if (launches) *launches = c->prof_sum[i].launches;
    if (ms) *ms = c->prof_sum[i].ms;
    if (bytes) *bytes = c->prof_sum[i].bytes;
    return CSA_GPU_OK;
}

// ---- download ---------------------------------------------------------------------------------------
extern "C" int csa_gpu_batch_download(csa_gpu_ctx *c, int *rotations, csa_gpu_set_info *info) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (!c->ran) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_batch_download before csa_gpu_batch_run");
    int nsets = c->nsets;
    Exec &ex = c->ex;
    c->h_set_flags.assign(nsets, 0); c->h_set_nchains.assign(nsets, 0); c->h_set_cyclic.assign(nsets, 0);
    TRY(d2h(ex, c->h_set_flags.data(), c->set_flags.p, sizeof(u32) * nsets));
    TRY(d2h(ex, c->h_set_nchains.data(), c->set_nchains.p, sizeof(u32) * nsets));
    TRY(d2h(ex, c->h_set_cyclic.data(), c->set_cyclic.p, sizeof(u32) * nsets));
    if (rotations) TRY(d2h(ex, rotations, c->rotations.p, sizeof(int) * (size_t)c->M));
    for (int s = 0; s < nsets; s++) {
        int status = CSA_SET_OK;
        u32 fl = c->h_set_flags[s];
        // in the order in which the reference gets there (csamsa.c:324 analyzeTree)
        if (fl & CSA_FLAG_UNDEFINED) status = CSA_SET_UNDEFINED;
        else if (c->h_set_nblocks[s] == 0) status = CSA_SET_NO_UNIQUE;
        else if (fl & CSA_FLAG_DEGENERATE) status = CSA_SET_DEGENERATE;
        else if (fl & CSA_FLAG_HANG) status = CSA_SET_NONTERMINATING;
        if (rotations && status != CSA_SET_OK)
            for (u32 k = c->h_set_seq0[s]; k < c->h_set_seq0[s + 1]; k++) rotations[k] = 0;
        if (info) {
            csa_gpu_set_info &o = info[s];
            o.status = status;
            o.nseqs = (int)(c->h_set_seq0[s + 1] - c->h_set_seq0[s]);
            o.count_collected = c->have_stats ? (int)c->h_set_collected[s] : -1;
            o.count_suffixfree = c->have_stats ? (int)c->h_set_suffixfree[s] : -1;
            o.count_unique = (int)c->h_set_nblocks[s];
            o.count_chains = (int)c->h_set_nchains[s];
            o.nblocks = (int)c->h_set_nblocks[s];
            o.chain_is_cyclic = (int)c->h_set_cyclic[s];
            o.block_offset = c->h_set_blk0[s];
        }
    }
    return CSA_GPU_OK;
}

extern "C" long long csa_gpu_batch_num_blocks(csa_gpu_ctx *c) { return (c && c->ran) ? (long long)c->B : -1; }
extern "C" long long csa_gpu_batch_num_positions(csa_gpu_ctx *c) { return (c && c->ran) ? (long long)c->E : -1; }

extern "C" int csa_gpu_batch_blocks(csa_gpu_ctx *c, int *depth, int *size, int *totalsize, int *interval, int *next,
                                    int *positions) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (!c->ran) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_batch_blocks before csa_gpu_batch_run");
    Exec &ex = c->ex;
    size_t nb = sizeof(int) * (size_t)c->B;
    if (c->B == 0) return CSA_GPU_OK;
    if (depth) TRY(d2h(ex, depth, c->f_depth.p, nb));
    if (size) TRY(d2h(ex, size, c->f_size.p, nb));
    if (totalsize) TRY(d2h(ex, totalsize, c->f_total.p, nb));
    if (interval) TRY(d2h(ex, interval, c->f_interval.p, nb));
    if (next) TRY(d2h(ex, next, c->f_next.p, nb));
    if (positions) TRY(d2h(ex, positions, c->f_pos.p, sizeof(int) * (size_t)c->E));
    return CSA_GPU_OK;
}

// the letters of every block, spelled as the reference's blockLabel spells them
extern "C" long long csa_gpu_batch_block_letters(csa_gpu_ctx *c, char *letters, long long *offsets) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (!c->ran) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_batch_block_letters before csa_gpu_batch_run");
    Exec &ex = c->ex;
    const u32 B = c->B;
    std::vector<int> depth(B);
    std::vector<unsigned long long> off((size_t)B + 1, 0);
    if (B) TRY(d2h(ex, depth.data(), c->f_depth.p, sizeof(int) * (size_t)B));
    for (u32 b = 0; b < B; b++) off[b + 1] = off[b] + (unsigned long long)depth[b];
    if (offsets) for (u32 b = 0; b <= B; b++) offsets[b] = (long long)off[b];
    const unsigned long long T = off[B];
    if (!letters || T == 0) return (long long)T;
#ifndef CSA_EMU
    CUDA_TRY(cudaSetDevice(c->device));
#endif
    TRY(dev_alloc(c->let_off, sizeof(unsigned long long) * ((size_t)B + 1)));
    TRY(dev_alloc(c->let_out, (size_t)T));
    TRY(h2d(ex, c->let_off.p, off.data(), sizeof(unsigned long long) * ((size_t)B + 1)));
    BlockLettersArgs a{view_of(c), P<unsigned char>(c->raw), c->q0,
                       P<int>(c->f_depth), P<int>(c->f_pos), P<u32>(c->f_leaf), P<u32>(c->f_set), P<u32>(c->set_blk0), P<u32>(c->set_pos0),
                       P<unsigned long long>(c->let_off), B, P<char>(c->let_out)};
    launch_blockletters(ex, (long long)T, a);
    TRY(d2h(ex, letters, c->let_out.p, (size_t)T));
    return (long long)T;
}

extern "C" int csa_gpu_batch_rotations(csa_gpu_ctx *c, int nsets, const int *set_start, const char *const *texts,
                                       const int *textsizes, int max_interval, unsigned flags, int *rotations,
                                       csa_gpu_set_info *info) {
    TRY(csa_gpu_batch_upload(c, nsets, set_start, texts, textsizes));
    TRY(csa_gpu_batch_run(c, max_interval, flags));
    return csa_gpu_batch_download(c, rotations, info);
}

extern "C" int csa_gpu_find_rotations(csa_gpu_ctx *c, int numberofseqs, const char *const *texts, const int *textsizes,
                                      int max_interval, unsigned flags, int *rotations, csa_gpu_set_info *info) {
    int set_start[2] = {0, numberofseqs};
    return csa_gpu_batch_rotations(c, 1, set_start, texts, textsizes, max_interval, flags, rotations, info);
}

// ---- introspection -----------------------------------------------------------------------------------
extern "C" long long csa_gpu_batch_num_suffixes(csa_gpu_ctx *c) { return (c && c->uploaded) ? (long long)c->N : -1; }

extern "C" int csa_gpu_batch_suffix_array(csa_gpu_ctx *c, unsigned *sa, int *lcp) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (!c->ran) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_batch_suffix_array before csa_gpu_batch_run");
    if (sa) TRY(d2h(c->ex, sa, c->sa.p, sizeof(u32) * (size_t)c->N));
    if (lcp) TRY(d2h(c->ex, lcp, c->t5.p, sizeof(u32) * (size_t)c->N));
    return CSA_GPU_OK;
}

extern "C" int csa_gpu_batch_timings(csa_gpu_ctx *c, float ms[6], long long *launches) {
    if (!c) CSA_FAIL(CSA_GPU_EINVAL, "null context");
    if (!c->ran) CSA_FAIL(CSA_GPU_ESTATE, "csa_gpu_batch_timings before csa_gpu_batch_run");
    if (ms) for (int i = 0; i < 6; i++) ms[i] = c->tm.ms[i];
    if (launches) *launches = c->launches;
    return CSA_GPU_OK;
}
