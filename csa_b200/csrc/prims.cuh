// prims.cuh -- hand-written device-wide primitives for sm_100a: prefix scans (sum / max over
// u32) and a stable LSD radix sort of (u64 key, u32 value) pairs.  No CUB/Thrust.
//
// Radix sort pass (8-bit digit), HBM traffic per element: histogram reads the key (8 B), the
// scatter reads key+value (12 B) and writes them (12 B) = 32 B.  Local ranks inside a 2048-key
// tile come from warp match-any ballots (stable), per-warp digit counters live in shared memory.
#pragma once
#include "csa_common.cuh"

struct ScanSum { HDM u32 id() { return 0u; } HDM u32 op(u32 a, u32 b) { return a + b; } };
struct ScanMax { HDM u32 id() { return 0u; } HDM u32 op(u32 a, u32 b) { return a > b ? a : b; } };

// scratch owned by the context; grown on demand
struct PrimScratch {
    DevMem block_sums[4]; // scan recursion levels
    DevMem counts;        // radix digit counters [256][nblocks]
    DevMem chain;         // single-pass scan: tile counter + one status word per tile
};

#define RS_TILE_ELEMS 2048 // keys per tile-block of a radix pass
// Segmented sort: tile-blocks never straddle a segment (a sequence set) and the digit counters are laid
// out segment by segment, so one exclusive scan gives offsets that keep every element inside its own
// segment -- the set number, the most significant part of the order, costs no radix pass.
struct RsSeg {
    const u32 *blk_start;  // [nblocks] first element of tile-block b
    const u32 *blk_count;  // [nblocks] elements in it (<= RS_TILE)
    const u32 *blk_cbase;  // [nblocks] index of the block's digit-0 counter
    const u32 *blk_stride; // [nblocks] distance between its counters of consecutive digits (blocks of its segment)
    u32 nblocks;
    const u32 *seg_bounds; // [nsegs+1] first element of every segment
    u32 nsegs;
};

#ifdef CSA_EMU
// ------------------------------- CPU emulation (tests only) -------------------------------
template <class Op, bool INCLUSIVE>
static int scan_u32(Exec &, PrimScratch &, const u32 *in, u32 *out, long long n, int = 0) {
    Op o;
    u32 run = o.id();
    for (long long i = 0; i < n; i++) {
        u32 v = in[i];
        if (INCLUSIVE) { run = o.op(run, v); out[i] = run; }
        else { out[i] = run; run = o.op(run, v); }
    }
    return 0;
}

template <class K>
static int radix_sort_pairs(Exec &, PrimScratch &, K *&keys, u32 *&vals, K *&keys_alt, u32 *&vals_alt,
                            long long n, int begin_bit, int end_bit, const RsSeg *seg = nullptr, bool iota_vals = false) {
    if (iota_vals) for (long long i = 0; i < n; i++) vals[i] = (u32)i; // (the values are the places themselves)
    if (n <= 1 || end_bit <= begin_bit) return 0;
    const int kb = (int)sizeof(K) * 8;
    K mask = (end_bit - begin_bit >= kb) ? (K)~(K)0 : (K)((((K)1 << (end_bit - begin_bit)) - 1) << begin_bit);
    std::vector<long long> idx(n);
    for (long long i = 0; i < n; i++) idx[i] = i;
    u32 one[2] = {0, (u32)n};
    const u32 *bounds = seg ? seg->seg_bounds : one;
    u32 nsegs = seg ? seg->nsegs : 1;
    for (u32 g = 0; g < nsegs; g++) // every segment on its own: elements never leave their segment
        std::stable_sort(idx.begin() + bounds[g], idx.begin() + bounds[g + 1],
                         [&](long long a, long long b) { return (keys[a] & mask) < (keys[b] & mask); });
    for (long long i = 0; i < n; i++) { keys_alt[i] = keys[idx[i]]; vals_alt[i] = vals[idx[i]]; }
    std::swap(keys, keys_alt);
    std::swap(vals, vals_alt);
    return 0;
}
#else
// ------------------------------------ CUDA ----------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <class Op>
__device__ __forceinline__ u32 warp_scan_incl(u32 v, Op o) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v = o.op(t, v);
    }
    return v;
}

// exclusive prefix of per-thread aggregates across the block; returns block total in `total`
template <class Op>
__device__ __forceinline__ u32 block_scan_excl(u32 agg, u32 &total, Op o, u32 *smem /*>=33*/) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = warp_scan_incl(agg, o);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 w = (lane < (SCAN_THREADS / 32)) ? smem[lane] : o.id();
        u32 wi = warp_scan_incl(w, o);
        smem[lane] = wi;
    }
    __syncthreads();
    u32 warp_prefix = warp ? smem[warp - 1] : o.id();
    total = smem[SCAN_THREADS / 32 - 1];
    u32 excl_in_warp = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl_in_warp = o.id();
    __syncthreads();
    return o.op(warp_prefix, excl_in_warp);
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const u32 *__restrict__ in, u32 *__restrict__ sums, long long n) {
    __shared__ u32 sm[33];
    Op o;
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    u32 agg = o.id();
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) agg = o.op(agg, in[base + j]);
    u32 total;
    block_scan_excl(agg, total, o, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const u32 *in, u32 *out, // in may alias out
                                                             const u32 *__restrict__ block_prefix, long long n) {
    __shared__ u32 sm[33];
    Op o;
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u32 agg = o.id();
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = (base + j < n) ? in[base + j] : o.id();
        agg = o.op(agg, v[j]);
    }
    u32 total;
    u32 excl = block_scan_excl(agg, total, o, sm);
    u32 run = block_prefix ? o.op(block_prefix[blockIdx.x], excl) : excl;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (INCLUSIVE) { run = o.op(run, v[j]); if (base + j < n) out[base + j] = run; }
        else { if (base + j < n) out[base + j] = run; run = o.op(run, v[j]); }
    }
}

// ---- single-pass scan (decoupled look-back) for long arrays: 4 B read + 4 B written per element ----
// Tiles take their number from an atomic counter, so a tile only ever waits for tiles that have
// already started.  A tile publishes (status, value) as ONE 64-bit word: status 1 = its own
// aggregate, 2 = the inclusive prefix up to and including it; warp 0 of a later tile walks back over
// those words, 32 at a time, until it meets an inclusive prefix.
#define CS_THREADS 256
#define CS_ITEMS 16
#define CS_TILE (CS_THREADS * CS_ITEMS)
template <class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(CS_THREADS) k_scan_chain(const u32 *in, u32 *out, long long n, unsigned long long *state) {
    __shared__ u32 sm[33];
    __shared__ u32 s_tile, s_prefix;
    Op o;
    if (threadIdx.x == 0) s_tile = (u32)atomicAdd(state, 1ull); // state[0]: next tile number; state[1+t]: tile t
    __syncthreads();
    const u32 tile = s_tile;
    volatile unsigned long long *st = state + 1;
    const long long base = (long long)tile * CS_TILE + (long long)threadIdx.x * CS_ITEMS;
    u32 v[CS_ITEMS];
    if (base + CS_ITEMS <= n && ((reinterpret_cast<size_t>(in) & 15) == 0)) {
        const uint4 *p4 = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
        for (int q = 0; q < CS_ITEMS / 4; q++) { uint4 t = p4[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
    } else {
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) v[j] = (base + j < n) ? in[base + j] : o.id();
    }
    u32 agg = o.id();
#pragma unroll
    for (int j = 0; j < CS_ITEMS; j++) agg = o.op(agg, v[j]);
    u32 total;
    u32 excl = block_scan_excl(agg, total, o, sm);
    if (threadIdx.x == 0) {
        st[tile] = ((unsigned long long)(tile == 0 ? 2u : 1u) << 32) | total;
        if (tile == 0) s_prefix = o.id();
    }
    if (tile > 0 && threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        u32 run = o.id();
        long long look = (long long)tile - 1; // the nearest tile not yet accounted for
        for (;;) {
            const long long t = look - lane;
            unsigned long long w = (t >= 0) ? st[t] : (2ull << 32); // before tile 0: an empty inclusive prefix
            while (__any_sync(0xffffffffu, (w >> 32) == 0)) w = (t >= 0) ? st[t] : (2ull << 32); // not published yet
            const unsigned incl = __ballot_sync(0xffffffffu, (w >> 32) == 2);
            const int stop = incl ? (__ffs((int)incl) - 1) : 32; // nearest inclusive prefix among these 32
            u32 part = ((int)lane <= stop && lane < 32) ? (u32)w : o.id();
            if ((int)lane > stop) part = o.id();
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) part = o.op(part, __shfl_xor_sync(0xffffffffu, part, d));
            run = o.op(part, run);
            if (incl) break;
            look -= 32;
        }
        if (lane == 0) {
            s_prefix = run;
            st[tile] = (2ull << 32) | o.op(run, total);
        }
    }
    __syncthreads();
    u32 r = o.op(s_prefix, excl);
#pragma unroll
    for (int j = 0; j < CS_ITEMS; j++) {
        if (INCLUSIVE) { r = o.op(r, v[j]); v[j] = r; }
        else { u32 t = v[j]; v[j] = r; r = o.op(r, t); }
    }
    if (base + CS_ITEMS <= n && ((reinterpret_cast<size_t>(out) & 15) == 0)) {
        uint4 *p4 = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
        for (int q = 0; q < CS_ITEMS / 4; q++) p4[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < CS_ITEMS; j++) if (base + j < n) out[base + j] = v[j];
    }
}

// out may alias in
template <class Op, bool INCLUSIVE>
static int scan_u32(Exec &ex, PrimScratch &ps, const u32 *in, u32 *out, long long n, int level = 0) {
    if (n <= 0) return 0;
    if (n > 4 * SCAN_TILE) { // more than a few tiles: one pass (decoupled look-back), one launch instead of three
        long long nt = (n + CS_TILE - 1) / CS_TILE;
        int rc = dev_alloc(ps.chain, sizeof(unsigned long long) * (size_t)(nt + 1));
        if (rc) return rc;
        CUDA_TRY(cudaMemsetAsync(ps.chain.p, 0, sizeof(unsigned long long) * (size_t)(nt + 1), ex.stream));
        PROF_BEGIN(ex, "k_scan_chain", 8.0 * n);
        k_scan_chain<Op, INCLUSIVE><<<(unsigned)nt, CS_THREADS, 0, ex.stream>>>(in, out, n, (unsigned long long *)ps.chain.p);
        PROF_END(ex);
        ex.launches++;
        return 0;
    }
    long long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        PROF_BEGIN(ex, "k_scan_apply", 8.0 * n);
        k_scan_apply<Op, INCLUSIVE><<<1, SCAN_THREADS, 0, ex.stream>>>(in, out, nullptr, n);
        PROF_END(ex);
        ex.launches++;
        return 0;
    }
    if (level >= 4) CSA_FAIL(-2, "scan too large");
    int rc = dev_alloc(ps.block_sums[level], sizeof(u32) * (size_t)nb);
    if (rc) return rc;
    u32 *sums = (u32 *)ps.block_sums[level].p;
    PROF_BEGIN(ex, "k_scan_reduce", 4.0 * n);
    k_scan_reduce<Op><<<(unsigned)nb, SCAN_THREADS, 0, ex.stream>>>(in, sums, n);
    PROF_END(ex);
    ex.launches++;
    rc = scan_u32<Op, false>(ex, ps, sums, sums, nb, level + 1);
    if (rc) return rc;
    PROF_BEGIN(ex, "k_scan_apply", 8.0 * n);
    k_scan_apply<Op, INCLUSIVE><<<(unsigned)nb, SCAN_THREADS, 0, ex.stream>>>(in, out, sums, n);
    PROF_END(ex);
    ex.launches++;
    return 0;
}

// ---- radix sort -----------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)
static_assert(RS_TILE == RS_TILE_ELEMS, "tile size");
#define RS_BITS 8
#define RS_BINS 256
#ifndef RS_HCOPIES
#define RS_HCOPIES 2
#endif

template <class K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K *__restrict__ keys, u32 *__restrict__ counts,
                                                        long long n, int shift, RsSeg seg) {
    // two copies of the counters (even and odd warps): a copy per warp cost more in zeroing and summing 2048 words a tile than
    // it saved in colliding atomics
    __shared__ u32 h[RS_HCOPIES][RS_BINS];
    for (int i = threadIdx.x; i < RS_HCOPIES * RS_BINS; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long start = seg.blk_start ? (long long)seg.blk_start[blockIdx.x] : (long long)blockIdx.x * RS_TILE;
    const long long end = seg.blk_start ? start + seg.blk_count[blockIdx.x] : (start + RS_TILE < n ? start + RS_TILE : n);
    long long base = start + (long long)warp * (RS_ITEMS * 32) + lane;
    K k[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) { const long long i = base + j * 32; k[j] = i < end ? keys[i] : (K)0; } // (all loads in flight before the first atomic)
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++)
        if (base + j * 32 < end) atomicAdd(&h[warp % RS_HCOPIES][(unsigned)(k[j] >> shift) & (RS_BINS - 1)], 1u);
    __syncthreads();
    const size_t cbase = seg.blk_start ? seg.blk_cbase[blockIdx.x] : blockIdx.x;
    const size_t stride = seg.blk_start ? seg.blk_stride[blockIdx.x] : seg.nblocks;
    for (int d = threadIdx.x; d < RS_BINS; d += RS_THREADS) {
        u32 s = 0;
#pragma unroll
        for (int w = 0; w < RS_HCOPIES; w++) s += h[w][d];
        counts[cbase + (size_t)d * stride] = s;
    }
}

template <class K>
__global__ void __launch_bounds__(RS_THREADS, 4) k_rs_scatter(const K *__restrict__ keys, const u32 *__restrict__ vals,
                                                           K *__restrict__ keys_out, u32 *__restrict__ vals_out,
                                                           const u32 *__restrict__ offsets, long long n, int shift, RsSeg seg) {
    __shared__ u32 h[RS_WARPS][RS_BINS];
    __shared__ u32 tstart[RS_BINS], gdelta[RS_BINS];
    __shared__ K s_keys[RS_TILE];
    __shared__ u32 s_vals[RS_TILE];
    __shared__ u32 s_scan[33];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const long long start = seg.blk_start ? (long long)seg.blk_start[blockIdx.x] : (long long)blockIdx.x * RS_TILE;
    const long long end = seg.blk_start ? start + seg.blk_count[blockIdx.x] : (start + RS_TILE < n ? start + RS_TILE : n);
    long long base = start + (long long)warp * (RS_ITEMS * 32) + lane;
    K k[RS_ITEMS];
    u32 v[RS_ITEMS];
    u32 rank[RS_ITEMS];
    unsigned dig[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) { // (the tile's loads are in flight while the counters are zeroed)
        long long i = base + j * 32;
        bool ok = i < end;
        k[j] = ok ? keys[i] : (K)0;
        v[j] = ok ? (vals ? vals[i] : (u32)i) : 0; // (vals == nullptr: the values are the places themselves, first pass of a sort)
    }
    const size_t cbase = seg.blk_start ? seg.blk_cbase[blockIdx.x] : blockIdx.x;
    const size_t stride = seg.blk_start ? seg.blk_stride[blockIdx.x] : seg.nblocks;
    const u32 goff = offsets[cbase + (size_t)threadIdx.x * stride]; // (where this tile's run of digit threadIdx.x starts in the output: asked for now, needed below)
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++)
        dig[j] = base + j * 32 < end ? ((unsigned)(k[j] >> shift) & (RS_BINS - 1)) : RS_BINS; // RS_BINS = "no key"
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        unsigned d = dig[j];
        unsigned peers = __match_any_sync(0xffffffffu, d);
        u32 before = 0;
        if (d < RS_BINS) before = h[warp][d];
        __syncwarp();
        if (d < RS_BINS && (peers & lt) == 0) h[warp][d] = before + __popc(peers);
        __syncwarp();
        rank[j] = before + __popc(peers & lt);
    }
    __syncthreads();
    // per digit: offsets of the warps inside the tile's run of that digit, the run's start inside the
    // digit-sorted tile (exclusive scan over the 256 digit totals) and its start in the output
    u32 tot = 0;
    {
        const int d = threadIdx.x; // RS_THREADS == RS_BINS
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            u32 t = h[w][d];
            h[w][d] = run;
            run += t;
        }
        tot = run;
    }
    u32 total;
    u32 dstart = block_scan_excl(tot, total, ScanSum(), s_scan);
    tstart[threadIdx.x] = dstart;
    // the element that ends up at place q of the sorted tile goes to out[gbase[d] + q - tstart[d]]:
    // fold both into one word so the write-out loop needs a single lookup
    gdelta[threadIdx.x] = goff - dstart;
    __syncthreads();
    // stage the tile in digit order in shared memory ...
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        unsigned d = dig[j];
        if (d < RS_BINS) {
            u32 q = tstart[d] + h[warp][d] + rank[j];
            s_keys[q] = k[j];
            s_vals[q] = v[j];
        }
    }
    __syncthreads();
    // ... and write it out in that order: neighbouring threads write neighbouring addresses inside
    // every digit's run (coalesced), instead of 32 scattered sectors per warp store
    const u32 cnt = (u32)(end - start);
    for (u32 q = threadIdx.x; q < cnt; q += RS_THREADS) {
        K kk = s_keys[q];
        unsigned d = (unsigned)(kk >> shift) & (RS_BINS - 1);
        u32 pos = gdelta[d] + q;
        keys_out[pos] = kk;
        vals_out[pos] = s_vals[q];
    }
}

// Stable LSD sort on key bits [begin_bit, end_bit), u32 or u64 keys, optionally segment by segment.
// keys/vals and the _alt buffers are swapped as passes go; on return keys/vals point at the sorted data.
template <class K>
static int radix_sort_pairs(Exec &ex, PrimScratch &ps, K *&keys, u32 *&vals, K *&keys_alt, u32 *&vals_alt,
                            long long n, int begin_bit, int end_bit, const RsSeg *segp = nullptr, bool iota_vals = false) {
    if (n <= 1 || end_bit <= begin_bit) {
        if (iota_vals) CSA_FAIL(-2, "radix sort: nothing to sort, values not made");
        return 0;
    }
    if (n >= (1ll << 32)) CSA_FAIL(-2, "radix sort: more than 2^32 elements");
    RsSeg seg;
    if (segp) seg = *segp;
    else { seg.blk_start = seg.blk_count = seg.blk_cbase = seg.blk_stride = seg.seg_bounds = nullptr; seg.nsegs = 1; seg.nblocks = (unsigned)((n + RS_TILE - 1) / RS_TILE); }
    unsigned nb = seg.nblocks;
    int rc = dev_alloc(ps.counts, sizeof(u32) * (size_t)RS_BINS * nb);
    if (rc) return rc;
    u32 *counts = (u32 *)ps.counts.p;
    const double kbytes = (double)sizeof(K);
    for (int shift = begin_bit; shift < end_bit; shift += RS_BITS) {
        PROF_BEGIN(ex, "k_rs_hist", kbytes * n);
        k_rs_hist<K><<<nb, RS_THREADS, 0, ex.stream>>>(keys, counts, n, shift, seg);
        PROF_END(ex);
        ex.launches++;
        rc = scan_u32<ScanSum, false>(ex, ps, counts, counts, (long long)RS_BINS * nb);
        if (rc) return rc;
        const bool iota = iota_vals && shift == begin_bit;
        PROF_BEGIN(ex, "k_rs_scatter", (2.0 * (kbytes + 4.0) - (iota ? 4.0 : 0.0)) * n);
        k_rs_scatter<K><<<nb, RS_THREADS, 0, ex.stream>>>(keys, iota ? nullptr : vals, keys_alt, vals_alt, counts, n, shift, seg);
        PROF_END(ex);
        ex.launches++;
        K *tk = keys; keys = keys_alt; keys_alt = tk;
        u32 *tv = vals; vals = vals_alt; vals_alt = tv;
    }
    return 0;
}
#endif
