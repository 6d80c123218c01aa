// cta_emu.h -- a CPU stand-in for ONE CUDA thread block at a time (TEST INFRASTRUCTURE ONLY).
//
// Kernels whose device code lives in a header free of CUDA-runtime host calls (e.g. csrc/plane_kernels.cuh) are compiled
// for the host with the macros below: every CUDA thread becomes a host thread, __syncthreads() a pthread barrier,
// __shared__ a function-local static (blocks run one after the other, so one copy is enough), atomics the GCC builtins.
// This checks the kernel's LOGIC -- indexing, phase structure, who writes what between barriers -- against the oracle on
// a machine without a GPU.  It says nothing about performance, memory-model subtleties or warp intrinsics.
#pragma once
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <time.h>
#include <stddef.h>
#include <stdint.h>

#include <thread>
#include <vector>

struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
};
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;
static pthread_barrier_t g_cta_barrier;

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __noinline__
#define __shared__ static

#define __align__(n) __attribute__((aligned(n)))

static inline void __syncthreads() { pthread_barrier_wait(&g_cta_barrier); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int atomicAdd(int *a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *a, unsigned v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int *a, int v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAnd(unsigned *a, unsigned v) { return __atomic_fetch_and(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int *a, int v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicCAS(int *a, int expected, int desired) {
    __atomic_compare_exchange_n(a, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;  // the value found, like CUDA's atomicCAS
}
static inline unsigned atomicAdd(unsigned *a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double *a, double v) {  // CAS loop on the bit pattern
    unsigned long long *p = reinterpret_cast<unsigned long long *>(a), old = __atomic_load_n(p, __ATOMIC_SEQ_CST), nw;
    double cur;
    do {
        __builtin_memcpy(&cur, &old, 8);
        const double sum = cur + v;
        __builtin_memcpy(&nw, &sum, 8);
    } while (!__atomic_compare_exchange_n(p, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return cur;
}
static inline unsigned long long atomicMax(unsigned long long *a, unsigned long long v) {
    unsigned long long old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T>
static inline T __ldcg(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
static inline double rsqrt(double v) { return 1.0 / sqrt(v); }
static inline long long __double_as_longlong(double v) {
    long long r;
    __builtin_memcpy(&r, &v, 8);
    return r;
}
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline long long clock64() {  // "cycles" = nanoseconds of the monotonic clock
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
}
struct float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct short2 {
    short x, y;
};
static inline int atomicMin(int *a, int v) {
    int old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}

// ---- vector types / scalar built-ins the kernels use
struct uint4 {
    unsigned x, y, z, w;
};
template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __float2int_rn(float v) { return (int)lrintf(v); }   // round-half-even (default rounding mode)
static inline int __double2int_rn(double v) { return (int)lrint(v); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

// ---- warp intrinsics: 32 consecutive host threads form a warp with its own barrier and exchange buffer.  Only
// full-mask, convergent use is supported (what the kernels do); a divergent shuffle deadlocks here -- it would be a bug
// on the GPU as well.
struct emu_warp_ctx {
    pthread_barrier_t bar;
    unsigned long long buf[32];
    unsigned lanes;
};
static thread_local emu_warp_ctx *emu_warp = nullptr;
static thread_local unsigned emu_lane = 0;

static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu_warp->bar); }
template <class T>
static inline T emu_exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    unsigned long long raw = 0;
    __builtin_memcpy(&raw, &v, sizeof(T));
    emu_warp->buf[emu_lane] = raw;
    pthread_barrier_wait(&emu_warp->bar);
    T out = v;
    if (src >= 0 && src < (int)emu_warp->lanes) {
        raw = emu_warp->buf[src];
        __builtin_memcpy(&out, &raw, sizeof(T));
    }
    pthread_barrier_wait(&emu_warp->bar);
    return out;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_exchange(v, (int)(emu_lane ^ (unsigned)o)); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int o) { return emu_exchange(v, (int)emu_lane + o < 32 ? (int)emu_lane + o : -1); }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int o) { return emu_exchange(v, (int)emu_lane - o); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu_warp->buf[emu_lane] = pred ? 1ull : 0ull;
    pthread_barrier_wait(&emu_warp->bar);
    unsigned m = 0;
    for (unsigned l = 0; l < emu_warp->lanes; ++l) m |= (emu_warp->buf[l] ? 1u : 0u) << l;
    pthread_barrier_wait(&emu_warp->bar);
    return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline unsigned __match_any_sync(unsigned, int key) {
    emu_warp->buf[emu_lane] = (unsigned long long)(unsigned)key;
    pthread_barrier_wait(&emu_warp->bar);
    unsigned m = 0;
    for (unsigned l = 0; l < emu_warp->lanes; ++l) m |= (emu_warp->buf[l] == (unsigned long long)(unsigned)key ? 1u : 0u) << l;
    pthread_barrier_wait(&emu_warp->bar);
    return m;
}

// ---- dynamic shared memory (kernels declare it with PLP_DYNAMIC_SMEM, csrc/devmath.cuh)
#define PLP_CTA_EMU 1
static uint8_t *emu_dynamic_smem = nullptr;

template <class Kernel, class... Args>
static void emu_launch2(Kernel kernel, unsigned grid_x, unsigned grid_y, unsigned block, size_t smem_bytes, Args... args);

template <class Kernel, class... Args>
static void emu_launch(Kernel kernel, unsigned grid, unsigned block, Args... args) {
    emu_launch2(kernel, grid, 1u, block, (size_t)0, args...);
}

// <<<(grid_x, grid_y), block, smem_bytes>>>
template <class Kernel, class... Args>
static void emu_launch2(Kernel kernel, unsigned grid_x, unsigned grid_y, unsigned block, size_t smem_bytes, Args... args) {
    gridDim.x = grid_x;
    gridDim.y = grid_y;
    blockDim.x = block;
    const unsigned nwarps = (block + 31) / 32;
    std::vector<emu_warp_ctx> warps(nwarps);
    std::vector<unsigned long long> smem(smem_bytes / 8 + 4);
    emu_dynamic_smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (unsigned bb = 0; bb < grid_x * grid_y; ++bb) {
        const unsigned b = bb % grid_x, by = bb / grid_x;
        pthread_barrier_init(&g_cta_barrier, nullptr, block);
        for (unsigned w = 0; w < nwarps; ++w) {
            warps[w].lanes = (w + 1) * 32 <= block ? 32 : block - w * 32;
            pthread_barrier_init(&warps[w].bar, nullptr, warps[w].lanes);
        }
        std::vector<std::thread> threads;
        threads.reserve(block);
        emu_warp_ctx *wp = warps.data();
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([=]() {
                threadIdx.x = t;
                blockIdx.x = b;
                blockIdx.y = by;
                emu_warp = wp + t / 32;
                emu_lane = t % 32;
                kernel(args...);
            });
        for (auto &th : threads) th.join();
        for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_destroy(&warps[w].bar);
        pthread_barrier_destroy(&g_cta_barrier);
    }
}
