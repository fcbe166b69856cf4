// Thin runtime layer: device memory, copies, launches.  CUDA in the product build; tests/hostsim compiles the same
// engine with -DROFL_EMUL against tests/hostsim/cuda_emul.h to execute it on CPU threads (test tool only).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#ifdef ROFL_EMUL
#include "cuda_emul.h"
inline void rt_check(int, const char *) {}
inline void *rt_malloc(size_t n, cudaStream_t) { void *p = aligned_alloc(256, (n + 255) / 256 * 256); if (!p) throw std::runtime_error("alloc"); return p; }
inline void rt_free(void *p, cudaStream_t) { free(p); }
inline void rt_h2d(void *d, const void *h, size_t n, cudaStream_t) { memcpy(d, h, n); }
inline void rt_d2h(void *h, const void *d, size_t n, cudaStream_t) { memcpy(h, d, n); }
inline void rt_d2d(void *d, const void *s, size_t n, cudaStream_t) { memcpy(d, s, n); }
inline void rt_memset(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); }
inline void rt_sync(cudaStream_t) {}
inline size_t rt_free_mem() { return (size_t)8 << 30; }
inline void *rt_host_alloc(size_t n) { return malloc(n ? n : 1); }
inline void rt_set_device(int) {}
inline void rt_host_free(void *p) { free(p); }
struct rt_event { };
inline void *rt_prof_begin(int, cudaStream_t) { return nullptr; }
inline void rt_prof_end(int, void *, cudaStream_t) {}
inline void rt_prof_work(int, double) {}
#else
#include <cuda_runtime.h>
inline void rt_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
void rt_count_launch(const char *name);
inline void *rt_malloc(size_t n, cudaStream_t s) { void *p = nullptr; rt_check(cudaMallocAsync(&p, n ? n : 16, s), "cudaMallocAsync"); return p; }
inline void rt_free(void *p, cudaStream_t s) { if (p) cudaFreeAsync(p, s); }
inline void rt_h2d(void *d, const void *h, size_t n, cudaStream_t s) { rt_check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "h2d"); }
inline void rt_d2h(void *h, const void *d, size_t n, cudaStream_t s) { rt_check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "d2h"); }
inline void rt_d2d(void *d, const void *s_, size_t n, cudaStream_t s) { rt_check(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s), "d2d"); }
inline void rt_memset(void *d, int v, size_t n, cudaStream_t s) { rt_check(cudaMemsetAsync(d, v, n, s), "memset"); }
inline void rt_sync(cudaStream_t s) { rt_check(cudaStreamSynchronize(s), "sync"); }
inline size_t rt_free_mem() { size_t f = 0, t = 0; cudaMemGetInfo(&f, &t); return f; }
inline void rt_set_device(int d) { rt_check(cudaSetDevice(d), "cudaSetDevice"); }     // the current device is per host thread
inline void *rt_host_alloc(size_t n) { void *p = nullptr; rt_check(cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault), "cudaHostAlloc"); return p; }     // pinned: async copies really are async
inline void rt_host_free(void *p) { if (p) cudaFreeHost(p); }
void *rt_prof_begin(int slot, cudaStream_t s);
void rt_prof_end(int slot, void *token, cudaStream_t s);
void rt_prof_work(int slot, double units);        // algorithmic work of the launches in that slot (e.g. point additions)
#endif

#define LAUNCH(k, grid, block, stream, ...) launch_##k(grid, block, stream, __VA_ARGS__)
#define LAUNCH_COOP LAUNCH

// stream-ordered scratch allocation with scope lifetime
struct dev_buf {
    void *p = nullptr; cudaStream_t s;
    dev_buf(size_t n, cudaStream_t st) : s(st) { p = rt_malloc(n, st); }
    ~dev_buf() { rt_free(p, s); }
    dev_buf(const dev_buf &) = delete; dev_buf &operator=(const dev_buf &) = delete;
    template <class T> T *as() const { return (T *)p; }
};
