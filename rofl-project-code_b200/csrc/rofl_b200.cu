// librofl_b200.so: the product library (CUDA, sm_100a only).  Context lifetime, CUDA-event profiling hooks, and the
// C ABI of include/rofl_b200.h (capi.cuh).  There is no CPU path in this translation unit.
#include "capi.cuh"
#include <atomic>
#include <cstring>

static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct prof_pair { cudaEvent_t a, b; };
static std::vector<prof_pair> g_prof_events[PROF_SLOTS];
static double g_prof_ms[PROF_SLOTS];
static long g_prof_launch[PROF_SLOTS];
static double g_prof_work[PROF_SLOTS];
static std::atomic<long> g_launches{0};

void rt_count_launch(const char *) { g_launches++; }
// ---- ROFL_TIMELINE=1: start / end of every kernel on its stream (events), relative to the first launch after the last dump
struct tl_rec { const char *name; cudaStream_t s; unsigned blocks; cudaEvent_t a, b; double host_ms; };
static double g_tl_host0 = 0;
static inline double tl_now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static std::mutex g_tl_mu; static std::vector<tl_rec *> g_tl; static cudaEvent_t g_tl_base = nullptr;
static const bool g_tl_on = getenv("ROFL_TIMELINE") != nullptr;
void *rt_timeline_begin(const char *name, cudaStream_t s, unsigned blocks) {
    if (!g_tl_on) return nullptr;
    tl_rec *r = new tl_rec{name, s, blocks, nullptr, nullptr, tl_now_ms()};
    cudaEventCreate(&r->a); cudaEventCreate(&r->b);
    { std::lock_guard<std::mutex> lk(g_tl_mu); if (!g_tl_base) { cudaEventCreate(&g_tl_base); cudaEventRecord(g_tl_base, s); g_tl_host0 = r->host_ms; } g_tl.push_back(r); }
    cudaEventRecord(r->a, s);
    return r;
}
void rt_timeline_end(void *tok, cudaStream_t s) { if (tok) cudaEventRecord(((tl_rec *)tok)->b, s); }
extern "C" void rofl_timeline_dump(const char *path) {
    if (!g_tl_on) return;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_tl_mu);
    FILE *f = fopen(path, "a"); if (!f) return;
    fprintf(f, "# name stream blocks start_ms end_ms host_enqueue_ms\n");
    for (tl_rec *r : g_tl) { float t0 = 0, t1 = 0; cudaEventElapsedTime(&t0, g_tl_base, r->a); cudaEventElapsedTime(&t1, g_tl_base, r->b); fprintf(f, "%s %p %u %.3f %.3f %.3f\n", r->name, (void *)r->s, r->blocks, t0, t1, r->host_ms - g_tl_host0); cudaEventDestroy(r->a); cudaEventDestroy(r->b); delete r; }
    g_tl.clear(); if (g_tl_base) { cudaEventDestroy(g_tl_base); g_tl_base = nullptr; }
    fclose(f);
}
void *rt_prof_begin(int, cudaStream_t s) {
    if (!g_prof_on.load()) return nullptr;
    cudaEvent_t a; cudaEventCreate(&a); cudaEventRecord(a, s); return (void *)a;
}
void rt_prof_end(int slot, void *token, cudaStream_t s) {
    if (!token) return;
    cudaEvent_t b; cudaEventCreate(&b); cudaEventRecord(b, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_events[slot].push_back({(cudaEvent_t)token, b}); g_prof_launch[slot]++;
}
void rt_prof_work(int slot, double units) { if (!g_prof_on.load()) return; std::lock_guard<std::mutex> lk(g_prof_mu); g_prof_work[slot] += units; }
static void prof_collect() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int sl = 0; sl < PROF_SLOTS; sl++) {
        for (auto &p : g_prof_events[sl]) { cudaEventSynchronize(p.b); float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b); g_prof_ms[sl] += ms; cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        g_prof_events[sl].clear();
    }
}
extern "C" void rofl_prof_enable(int on) { g_prof_on = on; }
extern "C" void rofl_prof_reset(void) { prof_collect(); for (int i = 0; i < PROF_SLOTS; i++) { g_prof_ms[i] = 0; g_prof_launch[i] = 0; g_prof_work[i] = 0; } g_launches = 0; }
extern "C" double rofl_prof_work(int slot) { std::lock_guard<std::mutex> lk(g_prof_mu); return (slot >= 0 && slot < PROF_SLOTS) ? g_prof_work[slot] : 0.0; }
extern "C" double rofl_prof_ms(int slot) { prof_collect(); return (slot >= 0 && slot < PROF_SLOTS) ? g_prof_ms[slot] : 0.0; }
extern "C" long rofl_prof_launches(int slot) { if (slot < 0) return g_launches.load(); return slot < PROF_SLOTS ? g_prof_launch[slot] : 0; }
// IMAD.WIDE.U32 issue-rate probe (the roofline denominator of this path): 8 accumulator chains per thread whose multiplicands change every
// iteration (nothing for ptxas to hoist), all SMs, best of 5.  Returns multiply-adds per second.
__global__ void k_probe_imad_wide(uint64_t *out, uint32_t y, int iters) {
    uint64_t c[8]; for (int k = 0; k < 8; k++) c[k] = threadIdx.x * 8 + k + y;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = (uint64_t)(uint32_t)c[(k + 1) & 7] * (uint32_t)c[(k + 2) & 7] + c[k];
    }
    uint64_t r = 0; for (int k = 0; k < 8; k++) r ^= c[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// second pattern: the carry-chained form the field multiplication uses (mad.lo.cc / madc.hi.cc -> IMAD.WIDE.U32 with carry-out)
__global__ void k_probe_imad_wide_cc(uint32_t *out, uint32_t y, int iters) {
    uint32_t a[8], t0 = threadIdx.x, t1 = 1, t2 = 2;
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 8 + k + y;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(t0), "+r"(t1), "+r"(t2) : "r"(a[k]), "r"(a[(k + 3) & 7]));
        a[i & 7] ^= t0;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = t0 ^ t1 ^ t2;
}
extern "C" double rofl_probe_imad_wide(rofl_ctx *c) {
    if (!c) return 0.0;
    try {
        rt_set_device(c->e.device);
        lane_guard lg(c->e); cudaStream_t s = lg.s();
        cudaDeviceProp prop; rt_check(cudaGetDeviceProperties(&prop, c->e.device), "props");
        const int tpb = 256, blocks = prop.multiProcessorCount * 32, iters = 4096;
        dev_buf buf(sizeof(uint64_t) * (size_t)tpb * blocks, s);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e30f;
        for (int r = 0; r < 12; r++) {          // both patterns, best of 5 each (first run of each = warm-up)
            cudaEventRecord(a, s);
            if (r < 6) k_probe_imad_wide<<<blocks, tpb, 0, s>>>(buf.as<uint64_t>(), 12345u + r, iters);
            else k_probe_imad_wide_cc<<<blocks, tpb, 0, s>>>(buf.as<uint32_t>(), 12345u + r, iters);
            cudaEventRecord(b, s);
            rt_check(cudaEventSynchronize(b), "probe"); float ms = 0; cudaEventElapsedTime(&ms, a, b); if (r % 6 && ms < best) best = ms;
        }
        cudaEventDestroy(a); cudaEventDestroy(b);
        return (double)tpb * blocks * iters * 8 / (best * 1e-3);
    } catch (const std::exception &ex) { g_last_error = ex.what(); return 0.0; }
}
extern "C" void *rofl_ctx_stream(rofl_ctx *c) { return (c && !c->e.lanes.empty()) ? (void *)c->e.lanes[0]->q[0].hi : nullptr; }      // lane 0: the one a single-threaded caller always gets


// Diagnostic (ROFL_JITTER=1): a busy host thread that records the largest gap between two of its own clock readings per 100 ms
// window (CLOCK_MONOTONIC, comparable with Python's time.monotonic()): a gap of milliseconds means the (virtual) CPU was taken away.
#include <chrono>
static std::atomic<bool> g_jit_stop{false};
static std::thread *g_jit_thread = nullptr;
static void jitter_start() {
    if (g_jit_thread || !getenv("ROFL_JITTER")) return;
    g_jit_thread = new std::thread([] {
        using clk = std::chrono::steady_clock;
        auto ms = [] { return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count(); };
        double last = ms(), wstart = last, wmax = 0, hstart = last, hmax = 0;
        fprintf(stderr, "[rofl jitter] probe thread running\n");
        while (!g_jit_stop.load(std::memory_order_relaxed)) {
            double t = ms(); if (t - last > wmax) wmax = t - last; last = t;
            if (t - wstart >= 100.0) { if (wmax > 0.5) fprintf(stderr, "[rofl jitter] t=%.0f ms: clock gap %.2f ms\n", wstart, wmax); if (wmax > hmax) hmax = wmax; wstart = t; wmax = 0; }
            if (t - hstart >= 2000.0) { fprintf(stderr, "[rofl jitter] heartbeat t=%.0f ms: largest gap of the last 2 s %.3f ms\n", t, hmax); hstart = t; hmax = 0; }
        }
    });
}
extern "C" int rofl_ctx_create(rofl_ctx **out, int device) {
    if (!out) return ROFL_ERR_ARGS;
    *out = nullptr;
    try {
        int ndev = 0;
        rt_check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
        if (device < 0 || device >= ndev) throw std::runtime_error("rofl_b200 needs a CUDA device (no CPU fallback): device index out of range");
        rt_check(cudaSetDevice(device), "cudaSetDevice");
        cudaDeviceProp prop; rt_check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
        if (prop.major < 10) throw std::runtime_error("rofl_b200 is built for sm_100a only");
        // the point formulas call fe_mul / fe_sq as real functions: give every thread room for the deepest call chain
        { size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitStackSize); if (cur < 8192) rt_check(cudaDeviceSetLimit(cudaLimitStackSize, 8192), "cudaDeviceSetLimit(stack)"); }
        jitter_start();
        rofl_ctx *c = new rofl_ctx();
        c->e.device = device;
        unsigned hc = std::thread::hardware_concurrency(); c->e.host_threads = hc ? (int)std::min(hc, 32u) : 8;
        if (const char *gv = getenv("ROFL_GROUPS")) c->e.groups = std::max(1, std::min(ROFL_MAX_GROUPS, atoi(gv)));
        if (const char *gv = getenv("ROFL_RT_PER")) c->e.rt_per = std::max(0, std::min(64, atoi(gv)));                     // table-MSM terms per thread (0 = one wave)
        if (const char *gv = getenv("ROFL_SPLIT")) { double w[ROFL_MAX_GROUPS]; int k = sscanf(gv, "%lf,%lf,%lf,%lf", &w[0], &w[1], &w[2], &w[3]); for (int i = 0; i < k; i++) if (w[i] > 0) c->e.group_w[i] = w[i]; }      // relative group sizes
        if (const char *gv = getenv("ROFL_RT")) c->e.use_rt = atoi(gv);                                   // 0 disables the generator tables
        if (const char *gv = getenv("ROFL_UNFOLD")) c->e.rt_unfold = std::max(0, std::min(6, atoi(gv)));   // unfolded IPP rounds (RT path)
        if (const char *gv = getenv("ROFL_RT_BITS")) c->e.rt_bits = std::max(8, std::min(RT_MAX_BITS, atoi(gv)));              // generator-table radix
        if (const char *gv = getenv("ROFL_FRZ")) c->e.use_frz = atoi(gv) != 0;                                            // frozen-level middle rounds
        if (const char *gv = getenv("ROFL_TAIL")) c->e.tail_np = std::max(0, std::min(TAIL_MAX_F / 2, atoi(gv)));    // 0 disables the fused IPP tail
        if (const char *gv = getenv("ROFL_TAIL_NCTA")) { const int v = atoi(gv); c->e.tail_ncta = v >= 8 ? 8 : v >= 4 ? 4 : v >= 2 ? 2 : 1; }
        engine_init(c->e);
        *out = c;
        return ROFL_OK;
    } catch (const std::exception &ex) { g_last_error = ex.what(); return ROFL_ERR_CUDA; }
}
extern "C" void rofl_ctx_destroy(rofl_ctx *c) {
    if (!c) return;
    try { cudaSetDevice(c->e.device); engine_destroy(c->e); } catch (...) {}
    delete c;
}
