// Thin runtime layer: device memory, copies, launches.  CUDA in the product build; tests/hostsim compiles the same
// engine with -DROFL_EMUL against tests/hostsim/cuda_emul.h to execute it on CPU threads (test tool only).
#pragma once
#include <stdexcept>
#include <exception>
#include <string>
#include <vector>
#include <mutex>
#include <cstdlib>
#include <cstdio>
#include <cstring>
// Diagnostic (ROFL_HOSTPROF=1): where a proof's HOST time goes -- stream waits, kernel launches, copies, allocation -- per calling
// thread; prove_chunks prints one line per call.
#include <chrono>
struct rt_host_prof { double sync = 0, launch = 0, copy = 0, alloc = 0, par = 0, ser = 0; };
inline bool rt_hostprof_on() { static const bool on = getenv("ROFL_HOSTPROF") != nullptr; return on; }
inline rt_host_prof &rt_hostprof() { static thread_local rt_host_prof p; return p; }
struct rt_host_gap { std::chrono::steady_clock::time_point last_end; const char *last_label = nullptr; int ordinal = 0; };
inline rt_host_gap &rt_hostgap() { static thread_local rt_host_gap g; return g; }
struct rt_host_timer {
    double *slot; std::chrono::steady_clock::time_point t0; const char *label;
    explicit rt_host_timer(double rt_host_prof::*m, const char *lb = "call") : slot(rt_hostprof_on() ? &(rt_hostprof().*m) : nullptr), label(lb) {
        if (!slot) return;
        t0 = std::chrono::steady_clock::now();
        rt_host_gap &g = rt_hostgap();               // host time BETWEEN two runtime calls: report the long ones with their neighbours
        if (g.last_label) { const double gap = std::chrono::duration<double, std::milli>(t0 - g.last_end).count(); if (gap > 3.0) fprintf(stderr, "[rofl gap] %.2f ms of host time between call #%d %s and %s\n", gap, g.ordinal, g.last_label, label); }
        g.ordinal++;
    }
    ~rt_host_timer() {
        if (!slot) return;
        auto t1 = std::chrono::steady_clock::now(); const double dt = std::chrono::duration<double, std::milli>(t1 - t0).count(); *slot += dt;
        if (dt > 1.0 && slot != &rt_hostprof().sync && slot != &rt_hostprof().par && slot != &rt_hostprof().ser) fprintf(stderr, "[rofl long] %.2f ms inside %s (call #%d)\n", dt, label, rt_hostgap().ordinal);
        rt_host_gap &g = rt_hostgap(); g.last_end = t1; g.last_label = label;
    }
};
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
inline void rt_d2h_finish(bool = false) {}
inline size_t rt_free_mem() { return (size_t)8 << 30; }
inline void rt_stream_after(cudaStream_t, cudaStream_t) {}
inline cudaStream_t rt_stream_create(int) { return nullptr; }
inline void rt_stream_destroy(cudaStream_t) {}
inline void *rt_raw_malloc(size_t n) { return rt_malloc(n, nullptr); }
inline void rt_raw_free(void *p) { free(p); }
inline void rt_trim() {}
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
// ROFL_TIMELINE=1 (diagnostic): a CUDA event before and after EVERY launch, dumped per stream by rofl_timeline_dump()
void *rt_timeline_begin(const char *name, cudaStream_t s, unsigned blocks);
void rt_timeline_end(void *tok, cudaStream_t s);
// Scratch allocation: a process-wide cache of device blocks (cudaMalloc once, reused forever).  The CUDA stream-ordered pool was the
// first choice, but with two chunk groups allocating and freeing on two streams it made one stream wait for the other's long kernels
// (cross-stream reuse), and forbidding that reuse made the pool grow with real allocations in the middle of a proof.
// Here a block is preferably handed back to the stream that returned it (no waiting at all); a block of another stream is taken only
// if the event recorded at its return has completed, else a new block is allocated.  Sizes are rounded up to 256 B; a block may serve requests down
// to half its size.
struct rt_big_block { void *p; size_t n; int dev; cudaStream_t s; cudaEvent_t ev; };
struct rt_big_cache {
    std::mutex mu; std::vector<rt_big_block> free_list; std::vector<rt_big_block> live;
    void *take(size_t n, cudaStream_t s) {
        n = (n + 255) / 256 * 256; if (!n) n = 256;
        int dev = 0; cudaGetDevice(&dev);
        rt_big_block b{}; bool found = false;
        { std::lock_guard<std::mutex> lk(mu);
          // a block this very stream returned needs no ordering at all; a block of another stream is taken only when the work that
          // preceded its return has already finished (event complete): waiting for another group's stream stalled whole proofs for
          // tens of ms and, once started, the groups kept stealing each other's blocks.  Otherwise allocate: the cache grows until
          // every stream owns what it needs (a few steps), then no call allocates or waits any more.
          size_t best = free_list.size();
          for (int pass = 0; pass < 2 && best == free_list.size(); pass++)
              for (size_t i = 0; i < free_list.size(); i++) {
                  const rt_big_block &fb = free_list[i];
                  if (fb.dev != dev || fb.n < n || fb.n > 2 * n || (best != free_list.size() && fb.n >= free_list[best].n)) continue;
                  if (pass == 0 ? fb.s == s : (fb.s != s && cudaEventQuery(fb.ev) == cudaSuccess)) best = i;
              }
          if (best < free_list.size()) { b = free_list[best]; free_list.erase(free_list.begin() + best); found = true; } }
        static const bool trace = getenv("ROFL_ALLOC_TRACE") != nullptr;
        if (trace && (!found || b.s != s)) fprintf(stderr, "[rofl alloc] %s %zu bytes (block %zu) stream %p\n", found ? "cross-stream reuse" : "cudaMalloc", n, found ? b.n : n, (void *)s);
        if (found) { if (b.s != s) rt_check(cudaStreamWaitEvent(s, b.ev, 0), "cudaStreamWaitEvent"); }
        else if (cudaMalloc(&b.p, n) == cudaSuccess || (cudaGetLastError(), trim(), cudaMalloc(&b.p, n) == cudaSuccess)) { b.n = n; b.dev = dev; rt_check(cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming), "cudaEventCreate"); }
        else {                                                       // out of memory: now a wait on another stream's block is the lesser evil
            cudaGetLastError();
            std::lock_guard<std::mutex> lk(mu);
            size_t best = free_list.size();
            for (size_t i = 0; i < free_list.size(); i++)
                if (free_list[i].dev == dev && free_list[i].n >= n && (best == free_list.size() || free_list[i].n < free_list[best].n)) best = i;
            if (best == free_list.size()) throw std::runtime_error("cudaMalloc: out of memory");
            b = free_list[best]; free_list.erase(free_list.begin() + best);
            rt_check(cudaStreamWaitEvent(s, b.ev, 0), "cudaStreamWaitEvent");
        }
        b.s = s;
        std::lock_guard<std::mutex> lk(mu); live.push_back(b);
        return b.p;
    }
    // hand every cached free block of this device back to CUDA (before a large long-lived allocation measures the free memory, or when cudaMalloc fails)
    void trim() {
        int dev = 0; cudaGetDevice(&dev);
        std::vector<rt_big_block> out;
        { std::lock_guard<std::mutex> lk(mu);
          for (size_t i = 0; i < free_list.size();) { if (free_list[i].dev == dev) { out.push_back(free_list[i]); free_list.erase(free_list.begin() + i); } else i++; } }
        for (auto &b : out) { cudaEventSynchronize(b.ev); cudaFree(b.p); cudaEventDestroy(b.ev); }
    }
    bool give(void *p, cudaStream_t s) {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < live.size(); i++) if (live[i].p == p) {
            rt_big_block b = live[i]; live.erase(live.begin() + i);
            b.s = s; cudaEventRecord(b.ev, s); free_list.push_back(b); return true;
        }
        return false;
    }
};
inline rt_big_cache &rt_bigs() { static rt_big_cache c; return c; }
inline void *rt_malloc(size_t n, cudaStream_t s) { rt_host_timer t(&rt_host_prof::alloc, "rt_malloc"); return rt_bigs().take(n, s); }
inline void rt_free(void *p, cudaStream_t s) { rt_host_timer t(&rt_host_prof::alloc, "rt_free"); if (p && !rt_bigs().give(p, s)) cudaFreeAsync(p, s); }
inline void rt_h2d(void *d, const void *h, size_t n, cudaStream_t s) { rt_host_timer t(&rt_host_prof::copy, "h2d"); rt_check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "h2d"); }
// Device -> host results.  cudaMemcpyAsync into PAGEABLE memory blocks the calling thread until the stream gets there -- and, measured on
// B200 (driver 580), it does so holding a driver lock: while one chunk group's thread sat in the copy of its finished proofs, the other
// groups' threads were stuck inside cudaLaunchKernel / cudaMalloc / cudaEventQuery, their last kernels reached the GPU only when the first
// group's had finished, and one proof in three took 44 ms instead of 38 (profiles/r02_scheduling.md).  So results always land in a pinned
// staging block first (or directly in the caller's buffer when that is pinned) and are handed over by the rt_sync() of that stream.
struct rt_pin_pool {
    std::mutex mu; std::vector<std::pair<void *, size_t>> free_;
    void *take(size_t n, size_t &cap) {
        { std::lock_guard<std::mutex> lk(mu);
          size_t best = free_.size();
          for (size_t i = 0; i < free_.size(); i++) if (free_[i].second >= n && (best == free_.size() || free_[i].second < free_[best].second)) best = i;
          if (best < free_.size() && free_[best].second <= 4 * n + 4096) { void *p = free_[best].first; cap = free_[best].second; free_.erase(free_.begin() + best); return p; } }
        cap = n < 4096 ? 4096 : (n + 4095) / 4096 * 4096;
        void *p = nullptr; rt_check(cudaHostAlloc(&p, cap, cudaHostAllocPortable), "cudaHostAlloc"); return p;
    }
    void give(void *p, size_t cap) { std::lock_guard<std::mutex> lk(mu); free_.emplace_back(p, cap); }
    void trim() { std::vector<std::pair<void *, size_t>> out; { std::lock_guard<std::mutex> lk(mu); out.swap(free_); } for (auto &b : out) cudaFreeHost(b.first); }
};
inline rt_pin_pool &rt_pins() { static rt_pin_pool p; return p; }
struct rt_pending_d2h { cudaStream_t s; void *h, *stage; size_t n, cap; };
inline std::vector<rt_pending_d2h> &rt_pending() { static thread_local std::vector<rt_pending_d2h> v; return v; }
inline bool rt_host_is_pinned(const void *h) { cudaPointerAttributes a; if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return false; } return a.type == cudaMemoryTypeHost; }
inline void rt_d2h(void *h, const void *d, size_t n, cudaStream_t s) {
    rt_host_timer t(&rt_host_prof::copy, "d2h");
    if (!n) return;
    if (n >= 65536 && rt_host_is_pinned(h)) { rt_check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "d2h"); return; }
    size_t cap = 0; void *st = rt_pins().take(n, cap);
    rt_check(cudaMemcpyAsync(st, d, n, cudaMemcpyDeviceToHost, s), "d2h");
    rt_pending().push_back({s, h, st, n, cap});
}
// hand over the staged results of stream s (all streams: s == nullptr && all) -- the caller has just synchronised it
inline void rt_d2h_deliver(cudaStream_t s, bool all = false) {
    auto &v = rt_pending();
    for (size_t i = 0; i < v.size();) {
        if (all || v[i].s == s) { memcpy(v[i].h, v[i].stage, v[i].n); rt_pins().give(v[i].stage, v[i].cap); v.erase(v.begin() + i); } else i++;
    }
}
// end of an API call: nothing may stay staged (a result copy without a matching rt_sync would otherwise be lost silently)
// (while an exception unwinds the call, the destinations may be dead stack frames: the staged results are dropped, not delivered)
inline void rt_d2h_finish(bool discard = false) {
    auto &v = rt_pending(); if (v.empty()) return;
    for (auto &p : v) cudaStreamSynchronize(p.s);
    if (discard) { for (auto &p : v) rt_pins().give(p.stage, p.cap); v.clear(); }
    else rt_d2h_deliver(nullptr, true);
}
inline void rt_d2d(void *d, const void *s_, size_t n, cudaStream_t s) { rt_check(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s), "d2d"); }
inline void rt_memset(void *d, int v, size_t n, cudaStream_t s) { rt_check(cudaMemsetAsync(d, v, n, s), "memset"); }
// Stream wait.  ROFL_SYNC=spin polls cudaStreamQuery instead (a proof has ~20 host round trips; measured on B200: no gain over the
// blocking call, so blocking stays the default)
inline void rt_sync(cudaStream_t s) {
    rt_host_timer t(&rt_host_prof::sync, "sync");
    static const int spin = [] { const char *m = getenv("ROFL_SYNC"); return m && m[0] == 's' ? 1 : 0; }();
    if (!spin) { rt_check(cudaStreamSynchronize(s), "sync"); rt_d2h_deliver(s); return; }
    for (;;) {
        cudaError_t e = cudaStreamQuery(s);
        if (e == cudaSuccess) { rt_d2h_deliver(s); return; }
        if (e != cudaErrorNotReady) rt_check(e, "sync");
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}
inline size_t rt_free_mem() { size_t f = 0, t = 0; cudaMemGetInfo(&f, &t); return f; }
inline void rt_trim() { rt_bigs().trim(); rt_pins().trim(); }
// long-lived allocations (generator tables, BSGS tables): straight from / back to CUDA, never through the scratch cache
inline void *rt_raw_malloc(size_t n) { void *p = nullptr; if (cudaMalloc(&p, n ? n : 256) != cudaSuccess) { cudaGetLastError(); rt_trim(); rt_check(cudaMalloc(&p, n ? n : 256), "cudaMalloc"); } return p; }
inline void rt_raw_free(void *p) { if (p) cudaFree(p); }
// level 0 = the greatest priority of the device, 1, 2, .. = lower (clamped to the least): pending blocks of a higher-priority stream are
// dispatched before those of lower ones whenever a block slot frees up (no preemption of running blocks)
inline cudaStream_t rt_stream_create(int level) {
    int least = 0, greatest = 0; cudaDeviceGetStreamPriorityRange(&least, &greatest);          // numerically: greatest <= least
    int p = greatest + level; if (p > least) p = least;
    cudaStream_t s; rt_check(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, p), "cudaStreamCreateWithPriority"); return s;
}
inline void rt_stream_destroy(cudaStream_t s) { if (s) cudaStreamDestroy(s); }
// work queued on `waiter` from now on starts after everything queued on `producer` so far (fork / join of a side stream)
inline void rt_stream_after(cudaStream_t waiter, cudaStream_t producer) {
    if (waiter == producer) return;
    cudaEvent_t ev; rt_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
    rt_check(cudaEventRecord(ev, producer), "cudaEventRecord"); rt_check(cudaStreamWaitEvent(waiter, ev, 0), "cudaStreamWaitEvent");
    cudaEventDestroy(ev);
}
inline void rt_set_device(int d) { rt_check(cudaSetDevice(d), "cudaSetDevice"); }     // the current device is per host thread
inline void *rt_host_alloc(size_t n) { void *p = nullptr; rt_check(cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault), "cudaHostAlloc"); return p; }     // pinned: async copies really are async
inline void rt_host_free(void *p) { if (p) cudaFreeHost(p); }
void *rt_prof_begin(int slot, cudaStream_t s);
void rt_prof_end(int slot, void *token, cudaStream_t s);
void rt_prof_work(int slot, double units);        // algorithmic work of the launches in that slot (e.g. point additions)
#endif

#define LAUNCH(k, grid, block, stream, ...) launch_##k(grid, block, stream, __VA_ARGS__)
#define LAUNCH_COOP LAUNCH

// One logical in-order queue over two CUDA streams of different priority.  The prover's throughput kernels (table MSMs, catch-up, table
// builds: thousands of long-running blocks) go to `lo`, the short latency-bound kernels of the Fiat-Shamir chain to `hi`, so that the chain
// of one chunk group is dispatched ahead of the bulk work of the other groups instead of waiting for a block slot behind it (DESIGN.md
// section 3 "scheduling").  Every switch is an event dependency, so the order of the queue is the program order.
struct chain {
    cudaStream_t hi = nullptr, lo = nullptr; int cur = 0;
    chain() {}
    chain(cudaStream_t h, cudaStream_t l) : hi(h), lo(l ? l : h) {}
    cudaStream_t on(int big) { if (big != cur) { rt_stream_after(big ? lo : hi, cur ? lo : hi); cur = big; } return cur ? lo : hi; }
    cudaStream_t small() { return on(0); }
    cudaStream_t big() { return on(1); }
};
// stream-ordered scratch allocation with scope lifetime
struct dev_buf {
    void *p = nullptr; cudaStream_t s; chain *q = nullptr;
    dev_buf(size_t n, cudaStream_t st) : s(st) { p = rt_malloc(n, st); }
    dev_buf(size_t n, chain &c) : s(c.hi), q(&c) { p = rt_malloc(n, c.hi); }          // owned by the queue: returned on its `hi` stream after joining `lo`
    ~dev_buf() { rt_free(p, q ? q->small() : s); }
    dev_buf(const dev_buf &) = delete; dev_buf &operator=(const dev_buf &) = delete;
    template <class T> T *as() const { return (T *)p; }
};
