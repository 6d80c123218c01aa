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
#define __shared__ static

static inline void __syncthreads() { pthread_barrier_wait(&g_cta_barrier); }
static inline int atomicAdd(int *a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }

template <class Kernel, class... Args>
static void emu_launch(Kernel kernel, unsigned grid, unsigned block, Args... args) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        pthread_barrier_init(&g_cta_barrier, nullptr, block);
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([=]() {
                threadIdx.x = t;
                blockIdx.x = b;
                kernel(args...);
            });
        for (auto &th : threads) th.join();
        pthread_barrier_destroy(&g_cta_barrier);
    }
}
