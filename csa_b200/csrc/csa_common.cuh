// csa_common.cuh -- build-mode glue for the B200 rotation finder.
//
// The product is compiled by nvcc for sm_100a.  The same sources also compile with g++ when
// CSA_EMU is defined: every one-thread-per-item kernel body then runs in a plain loop and the
// cooperative primitives (sort, scan) are replaced by <algorithm>.  That emulation exists ONLY
// so tests/ can single-step the kernel logic on a box without a GPU (tests/emu/); it is never
// built into libcsa_gpu.so and nothing in csa_b200/ loads it.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <vector>

typedef uint32_t u32;
typedef uint64_t u64;

#ifdef CSA_EMU
#include <algorithm>
#include <vector>
#define HD static inline
#define HDM inline
#define CSA_CLZLL(x) __builtin_clzll(x)
#define CSA_CLZ(x) __builtin_clz(x)
template <class T> static inline T emu_atomic_add(T *p, T v) { T o = *p; *p = (T)(o + v); return o; }
template <class T> static inline T emu_atomic_min(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T emu_atomic_max(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
#define ATOMIC_ADD(p, v) emu_atomic_add(p, v)
#define COUNT_IF(p, flag) do { if (flag) (*(p))++; } while (0)
#define ATOMIC_MIN(p, v) emu_atomic_min(p, v)
#define ATOMIC_MAX(p, v) emu_atomic_max(p, v)
#define ATOMIC_OR(p, v) (*(p) |= (v))
#define LDG(p) (*(p))
typedef int csaStream_t;
#else
#include <cuda_runtime.h>
#define HD static __device__ __forceinline__
#define HDM __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define CSA_CLZLL(x) __clzll((long long)(x))
#define CSA_CLZ(x) __clz((int)(x))
#define LDG(p) __ldg(p)
#else
#define CSA_CLZLL(x) __builtin_clzll(x)
#define CSA_CLZ(x) __builtin_clz(x)
#define LDG(p) (*(p))
#endif
#define ATOMIC_ADD(p, v) atomicAdd(p, v)
// *p += number of threads of the warp whose flag is set: one atomic per warp (ballot + popc)
#define COUNT_IF(p, flag)                                                        \
    do {                                                                         \
        unsigned act__ = __activemask();                                         \
        unsigned b__ = __ballot_sync(act__, (flag));                             \
        if (b__ && (threadIdx.x & 31) == (unsigned)(__ffs(act__) - 1)) atomicAdd((p), (u32)__popc(b__)); \
    } while (0)
#define ATOMIC_MIN(p, v) atomicMin(p, v)
#define ATOMIC_MAX(p, v) atomicMax(p, v)
#define ATOMIC_OR(p, v) atomicOr(p, v)
typedef cudaStream_t csaStream_t;
#endif

// ---- error plumbing ---------------------------------------------------------------------
extern thread_local char g_csa_err[512];
#define CSA_FAIL(code, ...)                                   \
    do {                                                      \
        snprintf(g_csa_err, sizeof(g_csa_err), __VA_ARGS__);  \
        return (code);                                        \
    } while (0)

#ifndef CSA_EMU
#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            snprintf(g_csa_err, sizeof(g_csa_err), "%s failed: %s (%s:%d)", #expr,            \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                            \
            return (e__ == cudaErrorMemoryAllocation) ? -3 : -4;                              \
        }                                                                                     \
    } while (0)
#endif

// ---- execution context ------------------------------------------------------------------
// Optional per-kernel timing: when Exec::prof is set every launch is bracketed by two CUDA events
// on the launching stream; csa_gpu_profile_* reports, per kernel name, launches, device time and
// the ALGORITHMIC bytes the launches moved (items x the per-item figure in DESIGN.md).  bench.py
// turns it on for a separate pass after the timed steps (roofline.achieved), never inside them.
struct ProfRec {
    const char *name;
    double bytes;
#ifndef CSA_EMU
    cudaEvent_t a, b;
#endif
};
struct Profiler {
#ifdef CSA_EMU
    int unused;
#else
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    size_t pool_used = 0;
    cudaEvent_t get() {
        if (pool_used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[pool_used++];
    }
#endif
};
struct Exec {
    csaStream_t stream;
    long long launches; // kernels launched since the last reset (reported as gpu_launches)
    Profiler *prof;
};
#ifdef CSA_EMU
#define PROF_BEGIN(ex, nm, by) do { (void)(ex); } while (0)
#define PROF_END(ex) do { (void)(ex); } while (0)
#else
static inline bool csa_trace_on() { static int on = -1; if (on < 0) on = getenv("CSA_GPU_TRACE") ? 1 : 0; return on == 1; }
#define PROF_BEGIN(ex, nm, by)                                                    \
    do {                                                                          \
        if (csa_trace_on()) { cudaStreamSynchronize((ex).stream); fprintf(stderr, "[csa] %s\n", (nm)); fflush(stderr); } \
        if ((ex).prof) {                                                          \
            ProfRec r__;                                                          \
            r__.name = (nm); r__.bytes = (double)(by);                            \
            r__.a = (ex).prof->get(); r__.b = (ex).prof->get();                   \
            cudaEventRecord(r__.a, (ex).stream);                                  \
            (ex).prof->recs.push_back(r__);                                       \
        }                                                                         \
    } while (0)
#define PROF_END(ex)                                                              \
    do { if ((ex).prof) cudaEventRecord((ex).prof->recs.back().b, (ex).stream); } while (0)
#endif

// ---- device memory ----------------------------------------------------------------------
struct DevMem {
    void *p = nullptr;
    size_t bytes = 0;
};

static inline int dev_alloc(DevMem &m, size_t bytes) {
    if (m.p && m.bytes >= bytes) return 0;
    if (m.p) {
#ifdef CSA_EMU
        free(m.p);
#else
        cudaFree(m.p);
#endif
        m.p = nullptr;
        m.bytes = 0;
    }
    if (bytes == 0) bytes = 16;
    bytes = (bytes + 255) & ~(size_t)255;
#ifdef CSA_EMU
    m.p = calloc(1, bytes);
    if (!m.p) CSA_FAIL(-3, "host allocation of %zu bytes failed", bytes);
#else
    CUDA_TRY(cudaMalloc(&m.p, bytes));
#endif
    m.bytes = bytes;
    return 0;
}

static inline void dev_free(DevMem &m) {
    if (!m.p) return;
#ifdef CSA_EMU
    free(m.p);
#else
    cudaFree(m.p);
#endif
    m.p = nullptr;
    m.bytes = 0;
}

static inline int dev_zero(Exec &ex, void *p, size_t bytes) {
#ifdef CSA_EMU
    (void)ex;
    memset(p, 0, bytes);
#else
    CUDA_TRY(cudaMemsetAsync(p, 0, bytes, ex.stream));
#endif
    return 0;
}

static inline int dev_fill_ff(Exec &ex, void *p, size_t bytes) {
#ifdef CSA_EMU
    (void)ex;
    memset(p, 0xFF, bytes);
#else
    CUDA_TRY(cudaMemsetAsync(p, 0xFF, bytes, ex.stream));
#endif
    return 0;
}

static inline int h2d(Exec &ex, void *dst, const void *src, size_t bytes) {
#ifdef CSA_EMU
    (void)ex;
    memcpy(dst, src, bytes);
#else
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ex.stream));
#endif
    return 0;
}

static inline int d2h(Exec &ex, void *dst, const void *src, size_t bytes) {
#ifdef CSA_EMU
    (void)ex;
    memcpy(dst, src, bytes);
#else
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ex.stream));
    CUDA_TRY(cudaStreamSynchronize(ex.stream));
#endif
    return 0;
}

static inline int d2d(Exec &ex, void *dst, const void *src, size_t bytes) {
#ifdef CSA_EMU
    (void)ex;
    memmove(dst, src, bytes);
#else
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ex.stream));
#endif
    return 0;
}

static inline int exec_sync(Exec &ex) {
#ifdef CSA_EMU
    (void)ex;
#else
    CUDA_TRY(cudaStreamSynchronize(ex.stream));
#endif
    return 0;
}

// ---- one-thread-per-item kernels ----------------------------------------------------------
// MAP_KERNEL(name, Args) turns `name_body(long long i, const Args &a)` into a named __global__
// kernel k_name (so ncu lists it by name) plus launch_name(ex, n, args).
#ifdef CSA_EMU
#define MAP_KERNEL(name, Args, BYTES_PER_ITEM)                          \
    static inline void launch_##name(Exec &ex, long long n, Args a) {   \
        (void)ex;                                                       \
        for (long long i = 0; i < n; i++) name##_body(i, a);            \
    }
#define MAP_KERNEL_N(name, Args, BYTES_PER_ITEM) MAP_KERNEL(name, Args, BYTES_PER_ITEM)
#else
#define MAP_KERNEL(name, Args, BYTES_PER_ITEM)                                            \
    __global__ void __launch_bounds__(256) k_##name(long long n, Args a) {                \
        long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;                   \
        if (i < n) name##_body(i, a);                                                     \
    }                                                                                     \
    static inline void launch_##name(Exec &ex, long long n, Args a) {                     \
        if (n <= 0) return;                                                               \
        PROF_BEGIN(ex, "k_" #name, (double)n * (BYTES_PER_ITEM));                         \
        k_##name<<<(unsigned)((n + 255) / 256), 256, 0, ex.stream>>>(n, a);               \
        PROF_END(ex);                                                                     \
        ex.launches++;                                                                    \
    }
// The same with MAP_ITEMS elements a thread (a CTA's threads side by side on each), for the LIGHT kernels over all suffixes of
// a batch: a thread an element is 400 000 CTAs for 10^8 suffixes, and the rate at which CTAs start (~1.5 per ns) then bounds
// the kernel at ~0.26 ms whatever it reads (k_initkey 0.48 -> 0.38 ms = 0.81 of the copy peak).  Not for kernels whose
// elements are few and heavy (a block, a set): they lose their parallelism.
#define MAP_ITEMS 4
#define MAP_KERNEL_N(name, Args, BYTES_PER_ITEM)                                          \
    __global__ void __launch_bounds__(256) k_##name(long long n, Args a) {                \
        long long i = (long long)blockIdx.x * (256 * MAP_ITEMS) + threadIdx.x;            \
        _Pragma("unroll 1")                                                               \
        for (int j = 0; j < MAP_ITEMS && i < n; j++, i += 256) name##_body(i, a);         \
    }                                                                                     \
    static inline void launch_##name(Exec &ex, long long n, Args a) {                     \
        if (n <= 0) return;                                                               \
        PROF_BEGIN(ex, "k_" #name, (double)n * (BYTES_PER_ITEM));                         \
        k_##name<<<(unsigned)((n + 256 * MAP_ITEMS - 1) / (256 * MAP_ITEMS)), 256, 0, ex.stream>>>(n, a); \
        PROF_END(ex);                                                                     \
        ex.launches++;                                                                    \
    }
#endif
