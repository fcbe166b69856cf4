// TEST TOOL: a tiny CUDA-on-CPU emulation (threads + barriers) so that the product's kernels (kernels.cuh) and host
// orchestration (engine.cuh) can be executed in this GPU-less container and compared with the oracle byte for byte.
// Only tests/hostsim builds with -DROFL_EMUL include this file; the product library is CUDA-only.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include <barrier>
#include <functional>
#include <memory>
#include <mutex>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
inline thread_local std::barrier<> *emu_bar = nullptr;
#define __shared__ static
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
inline void __syncthreads() { if (emu_bar) emu_bar->arrive_and_wait(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAnd(int *p, int v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicMax(uint32_t *p, uint32_t v) { uint32_t o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };

// cooperative launch: one OS thread per CUDA thread of a block, blocks run one after another
inline std::mutex &emu_launch_mu() { static std::mutex m; return m; }
template <class F> void emu_launch(dim3 grid, dim3 block, bool coop, F body) {
    std::lock_guard<std::mutex> one_at_a_time(emu_launch_mu());      // gridDim / blockDim / `__shared__` statics are process-wide: launches of different host threads (chunk groups) take turns
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        if (!coop) {
            blockIdx = dim3(bx, by, bz); emu_bar = nullptr;
            for (unsigned t = 0; t < block.x; t++) { threadIdx = dim3(t, 0, 0); body(); }
        } else {
            std::barrier<> bar(block.x);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < block.x; t++) th.emplace_back([&, t] {
                blockIdx = dim3(bx, by, bz); threadIdx = dim3(t, 0, 0); emu_bar = &bar;
                body();
                bar.arrive_and_drop();
            });
            for (auto &x : th) x.join();
        }
    }
}
