// Host orchestration of the rofl_crypto prove / verify hot path on one GPU (C++; the reference's host side is Rust).
// Mirrors, function for function, the vector API that rofl_service links:
//   range_proof_vec::{create_rangeproof, verify_rangeproof}      range_proof_vec/mod.rs:16-102,149-191
//   l2_range_proof_vec::{create_rangeproof_l2, verify_rangeproof_l2}   l2_range_proof_vec/mod.rs:15-140,185-228
//   square_proof_vec::{create_l2rangeproof_vec_existing, verify_l2rangeproof_vec}   square_proof_vec/mod.rs:19-75,130-160
//   pedersen_ops::{commit_vec, add_rp_vec_vec, discrete_log_vec_table}, ElGamal R halves, bsgs32::BSGSTable
// All chunks of one update advance in lock step: one kernel launch per phase covers every chunk, the sequential
// Fiat-Shamir transcripts (Merlin) stay on the host, one small device<->host exchange per challenge.
#pragma once
#include "kernels.cuh"
#include "rt.cuh"
#include <map>
#include <mutex>
#include <thread>
#include <atomic>
#ifndef ROFL_EMUL
#include <sys/resource.h>
#endif
#include <algorithm>
#include <chrono>
#include <cstdio>

enum { DOM_RANGE_PROVE = 1, DOM_RANGE_VERIFY = 2, DOM_SQUARE = 3, DOM_L2_PROVE = 4, DOM_L2_VERIFY = 5, DOM_CRP = 6, DOM_RND_VEC = 7, DOM_RANDPROOF = 8, DOM_SQUARE_RAND = 9 };
enum { PROF_FOLD = 0, PROF_MSM = 1, PROF_COMMIT = 2, PROF_SQUARE = 3, PROF_RTMSM = 4, PROF_TAIL = 5, PROF_FRZ = 6, PROF_SLOTS = 8 };

struct gens_entry { int n = 0; int cap = 0; niels_st *G = nullptr, *H = nullptr;
                    int rt_cap = 0, rt_c = 8; niels_st *RTG = nullptr, *RTH = nullptr; };     // radix-2^rt_c tables (RT path)
struct bsgs_entry { unsigned long long *keys = nullptr; uint32_t *vals = nullptr; uint32_t cap = 0; uint64_t size = 0; };
struct rofl_engine {
    int device = 0;
    cudaStream_t stream = 0;
    niels_st *tabB = nullptr, *tabH = nullptr;
    uint8_t B32[32], H32[32];
    std::map<int, gens_entry> gens;
    std::map<std::pair<uint64_t, int>, bsgs_entry> bsgs;
    std::mutex mu;
    int host_threads = 8;
    int groups = 3;                       // chunk groups proved / verified concurrently on separate streams (hides per-round latency)
    std::vector<cudaStream_t> gstreams;   // gstreams[0] == stream
    int use_rt = 1;                       // 0: never build generator tables (generic Pippenger / fold path only)
    int rt_unfold = 4;                    // IPP rounds computed over the original generators before the catch-up fold
    int tail_np = 32;                     // IPP rounds with half-size <= tail_np run in the fused on-device tail kernel (0 = off)
    std::mutex pin_mu; std::vector<std::pair<void *, size_t>> pins;      // pool of pinned host blocks for the per-round exchanges
    int use_frz = 1;                      // middle IPP rounds over frozen generators with on-the-fly Straus tables (kernels.cuh K6c)
    std::atomic<size_t> free_hint{0};     // free device memory when the generator tables were last (re)built: cudaMemGetInfo is NOT for the hot path
    int rt_bits = RT_MAX_BITS;            // widest generator-table radix to try (8..11)
    double rt_mem_frac = 0.45;            // tables may take this fraction of the free device memory
};

// pinned host scratch with scope lifetime, recycled through the engine's pool (cudaHostAlloc is far too slow to call per proof)
struct pinned_buf {
    rofl_engine &e; void *p = nullptr; size_t n = 0;
    pinned_buf(rofl_engine &eng, size_t bytes) : e(eng) {
        { std::lock_guard<std::mutex> lk(e.pin_mu);
          for (size_t i = 0; i < e.pins.size(); i++) if (e.pins[i].second >= bytes) { p = e.pins[i].first; n = e.pins[i].second; e.pins.erase(e.pins.begin() + i); break; } }
        if (!p) { n = std::max<size_t>(bytes, 1 << 16); p = rt_host_alloc(n); }
    }
    ~pinned_buf() { std::lock_guard<std::mutex> lk(e.pin_mu); e.pins.emplace_back(p, n); }
    pinned_buf(const pinned_buf &) = delete; pinned_buf &operator=(const pinned_buf &) = delete;
    template <class T> T *as() const { return (T *)p; }
};

// ROFL_TRACE=1: wall-clock phase marks of prove_chunks on stderr (stream synchronised at every mark; debugging aid only)
struct phase_trace {
    bool on; cudaStream_t s; std::chrono::steady_clock::time_point t0; std::string out;
    phase_trace(cudaStream_t st) : on(getenv("ROFL_TRACE") != nullptr), s(st) { if (on) { rt_sync(s); t0 = std::chrono::steady_clock::now(); } }
    void mark(const char *what) { if (!on) return; rt_sync(s); auto t = std::chrono::steady_clock::now(); char b[96]; snprintf(b, sizeof b, " %s=%.2f", what, std::chrono::duration<double, std::milli>(t - t0).count()); out += b; t0 = t; }
    ~phase_trace() { if (on) fprintf(stderr, "[rofl trace]%s\n", out.c_str()); }
};

// ---- small host helpers ---------------------------------------------------------------------------------------------------
static inline size_t next_pow2_sz(size_t v) { if (v <= 1) return 1; size_t n = v - 1; while (n & (n - 1)) n &= n - 1; return n << 1; }   // range_proof_vec/mod.rs:237-246
static inline int ilog2_sz(size_t x) { int l = 0; while (((size_t)1 << l) < x) l++; return l; }
static inline bool fp_ok(int n_bits, int frac) { return (n_bits == 8 || n_bits == 16 || n_bits == 32 || n_bits == 64) && frac >= 0 && frac <= 12; }
static inline float clip_max_f(int range, int n_bits, int frac) {                  // conversion32.rs:56-60
    uint64_t raw = (range - 1 >= 64) ? ~0ULL : ((1ULL << (range - 1)) - 1);
    return fix_to_f32(raw & fix_max(n_bits), frac);
}
static inline float l2_clip_max_f(int range, int n_bits, int frac) {               // conversion32.rs:62-64
    uint64_t raw = (range >= 64) ? ~0ULL : ((1ULL << range) - 1);
    return fix_to_f32(raw & fix_max(n_bits), frac);
}
static inline void derive_key(uint8_t out[32], const uint8_t seed[32], uint32_t domain, uint64_t index) {
    uint8_t buf[44]; memcpy(buf, seed, 32);
    for (int i = 0; i < 4; i++) buf[32 + i] = (uint8_t)(domain >> (8 * i));
    for (int i = 0; i < 8; i++) buf[36 + i] = (uint8_t)(index >> (8 * i));
    sha3_256(out, buf, 44);
}
static inline void key_words(uint32_t w[8], const uint8_t k[32]) { for (int i = 0; i < 8; i++) w[i] = (uint32_t)k[4 * i] | ((uint32_t)k[4 * i + 1] << 8) | ((uint32_t)k[4 * i + 2] << 16) | ((uint32_t)k[4 * i + 3] << 24); }
static inline void sc_to_st(sc_st &o, const sc &s) { for (int i = 0; i < 8; i++) o.w[i] = s.v[i]; }
static inline void st_to_sc(sc &s, const sc_st &o) { for (int i = 0; i < 8; i++) s.v[i] = o.w[i]; }
template <class F> static void parallel_for(size_t n, int threads, F f) {
    rt_host_timer t_(&rt_host_prof::par, "parallel_for");
    if (n <= 1 || threads <= 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    size_t nt = std::min<size_t>(threads, n); std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back([=] { for (size_t i = t; i < n; i += nt) f(i); });
    for (auto &x : th) x.join();
}
// light per-chunk host work (a few Keccak-f per chunk): thread start-up would cost more than the work
template <class F> static void serial_for(size_t n, F f) { rt_host_timer t_(&rt_host_prof::ser, "serial_for"); for (size_t i = 0; i < n; i++) f(i); }
// Montgomery's trick: v[i] <- v[i]^-1 (all non-zero)
static inline void sc_batch_invert(std::vector<sc> &v) {
    size_t n = v.size(); if (!n) return;
    std::vector<sc> pre(n); sc acc; sc_from_u64(acc, 1);
    for (size_t i = 0; i < n; i++) { pre[i] = acc; sc_mul(acc, acc, v[i]); }
    sc inv; sc_invert_vartime(inv, acc);           // public Fiat-Shamir challenges only
    for (size_t i = n; i-- > 0;) { sc t; sc_mul(t, inv, pre[i]); sc_mul(inv, inv, v[i]); v[i] = t; }
}
static inline void ts_challenge_scalar(transcript &t, const char *label, sc &out) { uint8_t b[64]; transcript_challenge(t, label, b, 64); sc_from_bytes_wide(out, b); }
static inline void ts_append_sc(transcript &t, const char *label, const sc &s) { uint8_t b[32]; sc_tobytes(b, s); transcript_append(t, label, b, 32); }
static inline bool is_zero32(const uint8_t *b) { uint8_t z = 0; for (int i = 0; i < 32; i++) z |= b[i]; return z == 0; }
// pow2[b] = s^(2^b), b < 32
static inline void sc_pow2_table(sc_st *tab, const sc &s) { sc c = s; for (int b = 0; b < 32; b++) { sc_to_st(tab[b], c); sc_mul(c, c, c); } }

// ---- engine lifetime ------------------------------------------------------------------------------------------------------
static inline void engine_init(rofl_engine &e) {
    ge_p3 B, H; ge_base(B); ge_compress(e.B32, B);
    uint8_t h[64]; sha3_512(h, e.B32, 32); ge_from_uniform_bytes(H, h); ge_compress(e.H32, H);     // PedersenGens::default / el_gamal.rs:31-40
    cudaStream_t s = e.stream;
    if (e.gstreams.empty()) e.gstreams.push_back(e.stream);
    e.free_hint = rt_free_mem();
    e.tabB = (niels_st *)rt_malloc(sizeof(niels_st) * FB_WINDOWS * FB_ENTRIES, s);
    e.tabH = (niels_st *)rt_malloc(sizeof(niels_st) * FB_WINDOWS * FB_ENTRIES, s);
    dev_buf pts(64, s);
    rt_h2d(pts.as<uint8_t>(), e.B32, 32, s); rt_h2d(pts.as<uint8_t>() + 32, e.H32, 32, s);
    int nent = FB_WINDOWS * FB_ENTRIES;
    LAUNCH(k_fb_table_build, dim3((nent + 63) / 64), dim3(64), s, e.tabB, pts.as<uint8_t>());
    LAUNCH(k_fb_table_build, dim3((nent + 63) / 64), dim3(64), s, e.tabH, pts.as<uint8_t>() + 32);
    rt_sync(s);
}
static inline void engine_destroy(rofl_engine &e) {
    cudaStream_t s = e.stream;
    rt_sync(s);
    rt_free(e.tabB, s); rt_free(e.tabH, s);
    for (auto &g : e.gens) { rt_free(g.second.G, s); rt_free(g.second.H, s); rt_free(g.second.RTG, s); rt_free(g.second.RTH, s); }
    for (auto &b : e.bsgs) { rt_free(b.second.keys, s); rt_free(b.second.vals, s); }
    for (auto &pp : e.pins) rt_host_free(pp.first);
    e.pins.clear();
    e.gens.clear(); e.bsgs.clear();
    rt_sync(s);
}
// BulletproofGens::new(n, m): cached per n, capacity grows (the reference rebuilds them per chunk per call,
// range_proof_vec/mod.rs:126,201)
static inline gens_entry &engine_gens(rofl_engine &e, int n, int m) {
    gens_entry &g = e.gens[n];
    if (g.cap >= m) return g;
    cudaStream_t s = e.stream;
    int newcap = std::max(m, g.cap * 2);
    niels_st *G = (niels_st *)rt_malloc(sizeof(niels_st) * (size_t)n * newcap, s);
    niels_st *H = (niels_st *)rt_malloc(sizeof(niels_st) * (size_t)n * newcap, s);
    if (g.cap) { rt_d2d(G, g.G, sizeof(niels_st) * (size_t)n * g.cap, s); rt_d2d(H, g.H, sizeof(niels_st) * (size_t)n * g.cap, s); }
    int cnt = 2 * (newcap - g.cap);
    LAUNCH(k_gens_build, dim3((cnt + 63) / 64), dim3(64), s, G, H, n, g.cap, newcap);
    rt_sync(s);
    rt_free(g.G, s); rt_free(g.H, s);
    g.n = n; g.cap = newcap; g.G = G; g.H = H;
    return g;
}

// radix-2^c generator tables for the first m parties of the n-bit generators (c = 10, 9 or 8: the widest that fits the memory
// budget; wider = fewer additions per term); returns false when not even c = 8 fits
static inline bool engine_rt(rofl_engine &e, gens_entry &g, int n, int m, rt_tables &out) {
    if (!e.use_rt) return false;
    auto fill = [&](rt_tables &t, int c) { t.c = c; t.nw = msm_nw(c); t.B = 1 << (c - 1); msm_recode_const(t.K, c); };
    if (g.rt_cap >= m) { out.G = g.RTG; out.H = g.RTH; fill(out, g.rt_c); return true; }
    cudaStream_t s = e.stream;
    const size_t cnt = (size_t)n * m;
    rt_sync(s); rt_free(g.RTG, s); rt_free(g.RTH, s); g.RTG = g.RTH = nullptr; g.rt_cap = 0; rt_sync(s);
    const size_t have = rt_free_mem();
    int c = std::max(8, std::min(RT_MAX_BITS, e.rt_bits));
    for (; c >= 8; c--) {
        const size_t nw = msm_nw(c), B = (size_t)1 << (c - 1);
        if ((double)(2 * cnt * nw * B * sizeof(niels_st) + cnt * nw * sizeof(p3_st)) <= e.rt_mem_frac * (double)have) break;
    }
    if (c < 8) return false;
    rt_tables t; fill(t, c);
    const size_t bytes = cnt * rt_row_entries(t) * sizeof(niels_st);
    niels_st *RTG = (niels_st *)rt_malloc(bytes, s), *RTH = (niels_st *)rt_malloc(bytes, s);
    {
        dev_buf P(cnt * t.nw * sizeof(p3_st), s);
        const size_t rows = cnt * t.nw, thr = rows * (t.B / 16);
        for (int which = 0; which < 2; which++) {
            LAUNCH(k_rt_shifts, dim3((unsigned)((cnt + 127) / 128)), dim3(128), s, P.as<p3_st>(), which ? g.H : g.G, (uint32_t)cnt, t.c, t.nw);
            LAUNCH(k_rt_rows, dim3((unsigned)((thr + 127) / 128)), dim3(128), s, which ? RTH : RTG, P.as<p3_st>(), rows, t.B);
        }
        rt_sync(s);
    }
    g.RTG = RTG; g.RTH = RTH; g.rt_cap = m; g.rt_c = c; t.G = RTG; t.H = RTH; out = t;
    e.free_hint = rt_free_mem();
    return true;
}
// blocks per msm for the direct table MSM: every block of a wave runs ceil(T / (nb*128)) terms per thread, so pick the nb whose
// waves (148 SMs x 4 resident blocks) x terms-per-thread product is smallest (+ a block-sum epilogue worth ~half a term)
static inline int rt_blocks(size_t T, int C) {
    const size_t slots = 148 * RTM_BLOCKS, max_nb = std::max<size_t>(1, std::min((T + 127) / 128, slots * 4 / (size_t)C));
    size_t best = 1; double best_cost = 1e300;
    for (size_t nb = 1; nb <= max_nb; nb++) {
        const double waves = (double)((nb * C + slots - 1) / slots), per = (double)((T + nb * 128 - 1) / (nb * 128));
        const double cost = waves * (per + 0.5);
        if (cost < best_cost * 0.999) { best_cost = cost; best = nb; }
    }
    return (int)best;
}
static inline void run_rt_msm(rofl_engine &e, cudaStream_t s, rt_msm_args a, int nb, uint32_t n_msm) {
    rt_prof_work(PROF_RTMSM, (double)a.T * n_msm * a.rt.nw);          // mixed additions (upper bound: zero digits skip theirs)
    void *tk = rt_prof_begin(PROF_RTMSM, s);
    LAUNCH_COOP(k_rt_msm, dim3(nb, n_msm), dim3(128), s, a);
    rt_prof_end(PROF_RTMSM, tk, s);
}

// ---- MSM front end ----------------------------------------------------------------------------------------------------------
// window width and term slicing for an MSM of T terms run as n_msm independent instances (kernels.cuh, k_msm)
struct msm_plan { int c, nw; uint32_t slices, slice_len; size_t out_count(size_t n_msm) const { return n_msm * slices * (size_t)nw; } };
static inline int msm_pick_c(size_t T) {
    int best = 8; double bc = 1e300;
    for (int c = 3; c <= 8; c++) { const double B = (double)(1 << (c - 1)), cost = (double)msm_nw(c) * ((double)T + B * c); if (cost < bc) { bc = cost; best = c; } }
    return best;
}
static inline msm_plan msm_plan_for(size_t T, size_t n_msm) {
    msm_plan p; p.c = msm_pick_c(T); p.slices = 1; p.slice_len = (uint32_t)T;
    if (const char *fc = getenv("ROFL_MSM_C")) {       // test hook: force the window width / slicing
        p.c = std::max(3, std::min(8, atoi(fc)));
        if (const char *fs = getenv("ROFL_MSM_SLICES")) { p.slices = (uint32_t)std::max<size_t>(1, std::min<size_t>(atoi(fs), T)); p.slice_len = (uint32_t)((T + p.slices - 1) / p.slices); p.slices = (uint32_t)((T + p.slice_len - 1) / p.slice_len); }
        p.nw = msm_nw(p.c); return p;
    }
    const size_t wpb = MSM_THREADS >> (p.c - 1), blocks = n_msm * ((msm_nw(p.c) + wpb - 1) / wpb), slots = 148 * 4;
    if (blocks < slots && T >= 4096) {
        p.slices = (uint32_t)std::min<size_t>((slots + blocks - 1) / blocks, T / 2048);
        p.slice_len = (uint32_t)((T + p.slices - 1) / p.slices); p.slices = (uint32_t)((T + p.slice_len - 1) / p.slice_len);
        p.c = msm_pick_c(p.slice_len);
    }
    p.nw = msm_nw(p.c);
    return p;
}
// a: v[], split, nseg, T, scalar_stride, out filled by the caller
static inline void run_msm(rofl_engine &e, cudaStream_t s, msm_args a, const msm_plan &p, uint32_t n_msm) {
    a.c = p.c; a.nw = p.nw; a.slices = p.slices; a.slice_len = p.slice_len; msm_recode_const(a.K, p.c);
    const int wpb = MSM_THREADS >> (p.c - 1);
    void *tk = rt_prof_begin(PROF_MSM, s);
    LAUNCH_COOP(k_msm, dim3((p.nw + wpb - 1) / wpb, n_msm, p.slices), dim3(MSM_THREADS), s, a);
    rt_prof_end(PROF_MSM, tk, s);
}
static inline msm_seg mk_seg(const void *base, uint32_t count, uint32_t stride, int kind) { msm_seg s; s.base = base; s.count = count; s.stride = stride; s.kind = kind; return s; }
static inline void run_finalize(cudaStream_t s, const finalize_args &f) { LAUNCH_COOP(k_finalize, dim3(f.count), dim3(FIN_THREADS), s, f); }
static inline void fin_windows(finalize_args &f, const p3_st *win, const msm_plan &p) { f.windows = win; f.c = p.c; f.nw = p.nw; f.slices = p.slices; }

// =============================================================================================================================
// RangeProof::prove_multiple for C chunks in lock step.  All pointers are device pointers except
// h_proofs (host).  d_vals: C*m shifted values, d_blind: C*m blindings (reduced), d_V32: C*m compressed commitments.
// label = transcript label ("RangeProof" / "L2RangeProof"); keys = C ChaCha20 keys (host).
// =============================================================================================================================
static void prove_chunks(rofl_engine &e, cudaStream_t s, const char *label, int n, int m, int C, const gens_entry &g, const rt_tables *rt, const uint64_t *d_vals,
                         const sc_st *d_blind, const uint8_t *d_V32, const std::vector<uint8_t> &keys, uint8_t *h_proofs) {
    const size_t N = (size_t)n * m, NT = N * C;
    const int lgN = ilog2_sz(N);
    const size_t plen = 32 * (9 + 2 * (size_t)lgN);
    phase_trace tr(s);
#ifndef ROFL_EMUL
    struct host_prof_line {          // ROFL_HOSTPROF=1: one line per call with the host time of this thread by category
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(); rt_host_prof p0 = rt_hostprof(); int C; struct rusage r0;
        host_prof_line() { getrusage(RUSAGE_THREAD, &r0); }
        ~host_prof_line() {
            if (!rt_hostprof_on()) return;
            struct rusage r1; getrusage(RUSAGE_THREAD, &r1);
            auto tv = [](const timeval &a, const timeval &b) { return (a.tv_sec - b.tv_sec) * 1e3 + (a.tv_usec - b.tv_usec) / 1e3; };
            fprintf(stderr, "[rofl host] rusage: user=%.2f ms sys=%.2f ms minor_faults=%ld major_faults=%ld vol_cs=%ld invol_cs=%ld\n", tv(r1.ru_utime, r0.ru_utime), tv(r1.ru_stime, r0.ru_stime),
                    r1.ru_minflt - r0.ru_minflt, r1.ru_majflt - r0.ru_majflt, r1.ru_nvcsw - r0.ru_nvcsw, r1.ru_nivcsw - r0.ru_nivcsw);
            const rt_host_prof &p = rt_hostprof(); const double tot = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            const double sy = p.sync - p0.sync, la = p.launch - p0.launch, co = p.copy - p0.copy, al = p.alloc - p0.alloc, pa = p.par - p0.par, se = p.ser - p0.ser;
            fprintf(stderr, "[rofl host] prove_chunks C=%d total=%.2f sync=%.2f launch=%.2f copy=%.2f alloc=%.2f parallel_for=%.2f serial_for=%.2f other=%.2f ms\n", C, tot, sy, la, co, al, pa, se, tot - sy - la - co - al - pa - se);
        }
    } hpl; hpl.C = C;
#endif
    // ---- device scratch
    dev_buf d_keys(32 * (size_t)C, s), d_sLR(sizeof(sc_st) * 2 * NT, s), d_sums(sizeof(sc_st) * 5 * C, s);
    { std::vector<uint32_t> kw(8 * (size_t)C); for (int c = 0; c < C; c++) key_words(&kw[8 * c], &keys[32 * c]); rt_h2d(d_keys.p, kw.data(), 32 * (size_t)C, s); }
    LAUNCH(k_nonces, dim3((unsigned)((NT + 255) / 256)), dim3(256), s, d_sLR.as<sc_st>(), d_keys.as<uint32_t>(), n, m, NT);
    LAUNCH_COOP(k_party_sums, dim3(C), dim3(256), s, d_sums.as<sc_st>(), d_keys.as<uint32_t>(), d_blind, (const sc_st *)nullptr, n, m, 0);
    // ---- A
    const int nbA = (int)std::min<size_t>(64, (N + 511) / 512);
    dev_buf d_partA(sizeof(p3_st) * (size_t)C * nbA, s), d_AS(64 * (size_t)C, s);
    LAUNCH_COOP(k_bits_sum, dim3(nbA, C), dim3(128), s, d_partA.as<p3_st>(), d_vals, g.G, g.H, n, m);
    {
        finalize_args f = {}; f.partial = d_partA.as<p3_st>(); f.npartial = nbA; f.sHa = d_sums.as<sc_st>(); f.tabB = e.tabB; f.tabH = e.tabH;
        f.out32 = d_AS.as<uint8_t>(); f.count = C;
        run_finalize(s, f);
    }
    // ---- S = (sum s_bl) H + <s_L, G> + <s_R, H>
    if (rt) {
        const int nbS = rt_blocks(2 * N, C);
        dev_buf d_partS(sizeof(p3_st) * (size_t)C * nbS, s);
        rt_msm_args a = {}; a.scalars = d_sLR.as<sc_st>(); a.T = (uint32_t)(2 * N); a.scalar_stride = (uint32_t)(2 * N); a.nG = (uint32_t)N; a.mode = 0; a.rt = *rt; a.partial = d_partS.as<p3_st>();
        run_rt_msm(e, s, a, nbS, C);
        finalize_args f = {}; f.partial = d_partS.as<p3_st>(); f.npartial = nbS; f.sHa = d_sums.as<sc_st>() + C; f.tabB = e.tabB; f.tabH = e.tabH;
        f.out32 = d_AS.as<uint8_t>() + 32 * (size_t)C; f.count = C;
        run_finalize(s, f);
    } else {
        const msm_plan pl = msm_plan_for(2 * N, C);
        dev_buf d_winS(sizeof(p3_st) * pl.out_count(C), s);
        msm_args a = {}; a.v[0].scalars = d_sLR.as<sc_st>(); a.split = (uint32_t)C; a.T = (uint32_t)(2 * N); a.scalar_stride = (uint32_t)(2 * N);
        a.v[0].seg[0] = mk_seg(g.G, (uint32_t)N, 0, 0); a.v[0].seg[1] = mk_seg(g.H, (uint32_t)N, 0, 0); a.nseg = 2; a.out = d_winS.as<p3_st>();
        run_msm(e, s, a, pl, C);
        finalize_args f = {}; fin_windows(f, d_winS.as<p3_st>(), pl); f.sHa = d_sums.as<sc_st>() + C; f.tabB = e.tabB; f.tabH = e.tabH;
        f.out32 = d_AS.as<uint8_t>() + 32 * (size_t)C; f.count = C;
        run_finalize(s, f);
    }
    std::vector<uint8_t> hV(32 * (size_t)C * m), hAS(64 * (size_t)C);
    rt_d2h(hV.data(), d_V32, hV.size(), s); rt_d2h(hAS.data(), d_AS.p, hAS.size(), s);
    rt_sync(s);
    // ---- transcripts: V..., A, S -> y, z
    std::vector<transcript> ts(C);
    std::vector<sc> y(C), z(C), yinv(C);
    parallel_for(C, e.host_threads, [&](size_t c) {
        transcript &t = ts[c]; transcript_init(t, label);
        transcript_append(t, "dom-sep", (const uint8_t *)"rangeproof v1", 13);
        transcript_append_u64(t, "n", (uint64_t)n); transcript_append_u64(t, "m", (uint64_t)m);
        for (int j = 0; j < m; j++) transcript_append(t, "V", &hV[32 * (c * m + j)], 32);
        uint8_t *o = h_proofs + plen * c;
        memcpy(o, &hAS[32 * c], 32); memcpy(o + 32, &hAS[32 * (C + c)], 32);
        transcript_append(t, "A", o, 32); transcript_append(t, "S", o + 32, 32);
        ts_challenge_scalar(t, "y", y[c]); ts_challenge_scalar(t, "z", z[c]);
    });
    yinv = y; sc_batch_invert(yinv);
    std::vector<sc_st> h_ypow2(32 * (size_t)C), h_zpow2(32 * (size_t)C), h_yinvpow2(32 * (size_t)C), h_z(C);
    for (int c = 0; c < C; c++) { sc_pow2_table(&h_ypow2[32 * c], y[c]); sc_pow2_table(&h_zpow2[32 * c], z[c]); sc_pow2_table(&h_yinvpow2[32 * c], yinv[c]); sc_to_st(h_z[c], z[c]); }
    dev_buf d_ypow2(sizeof(sc_st) * 32 * C, s), d_zpow2(sizeof(sc_st) * 32 * C, s), d_yinvpow2(sizeof(sc_st) * 32 * C, s), d_z(sizeof(sc_st) * C, s);
    rt_h2d(d_ypow2.p, h_ypow2.data(), sizeof(sc_st) * 32 * C, s); rt_h2d(d_zpow2.p, h_zpow2.data(), sizeof(sc_st) * 32 * C, s);
    rt_h2d(d_yinvpow2.p, h_yinvpow2.data(), sizeof(sc_st) * 32 * C, s); rt_h2d(d_z.p, h_z.data(), sizeof(sc_st) * C, s);
    // ---- polynomials
    const int nbP = (int)std::min<size_t>(256, (N + 255) / 256);
    dev_buf d_a(sizeof(sc_st) * NT, s), d_b(sizeof(sc_st) * NT, s), d_part(sizeof(sc_st) * 3 * (size_t)C * nbP, s), d_tsum(sizeof(sc_st) * 3 * C, s);
    // split power tables for y^k, y^-k (k < N) and z^j (j < m)
    const int lgm = ilog2_sz((size_t)m);
    pow_tab ytab = {nullptr, lgN / 2, lgN - lgN / 2}, yitab = ytab, ztab = {nullptr, lgm / 2, lgm - lgm / 2};
    dev_buf d_ptab(sizeof(sc_st) * (size_t)C * (2 * pow_tab_size(ytab) + pow_tab_size(ztab)), s);
    ytab.tab = d_ptab.as<sc_st>(); yitab.tab = ytab.tab + (size_t)C * pow_tab_size(ytab); ztab.tab = yitab.tab + (size_t)C * pow_tab_size(yitab);
    LAUNCH(k_pow_tables, dim3((pow_tab_size(ytab) + 255) / 256, C), dim3(256), s, (sc_st *)ytab.tab, d_ypow2.as<sc_st>(), ytab.L, ytab.H);
    LAUNCH(k_pow_tables, dim3((pow_tab_size(yitab) + 255) / 256, C), dim3(256), s, (sc_st *)yitab.tab, d_yinvpow2.as<sc_st>(), yitab.L, yitab.H);
    LAUNCH(k_pow_tables, dim3((pow_tab_size(ztab) + 255) / 256, C), dim3(256), s, (sc_st *)ztab.tab, d_zpow2.as<sc_st>(), ztab.L, ztab.H);
    LAUNCH_COOP(k_poly, dim3(nbP, C), dim3(256), s, d_a.as<sc_st>(), d_b.as<sc_st>(), d_sLR.as<sc_st>(), d_part.as<sc_st>(), d_vals, ytab, ztab, d_zpow2.as<sc_st>(), n, m);
    LAUNCH_COOP(k_sc_sum, dim3(C), dim3(256), s, d_tsum.as<sc_st>(), d_part.as<sc_st>(), nbP, 3);
    LAUNCH_COOP(k_party_sums, dim3(C), dim3(256), s, d_sums.as<sc_st>(), d_keys.as<uint32_t>(), d_blind, d_z.as<sc_st>(), n, m, 1);
    std::vector<sc_st> h_tsum(3 * (size_t)C), h_sums(5 * (size_t)C);
    rt_d2h(h_tsum.data(), d_tsum.p, sizeof(sc_st) * 3 * C, s); rt_d2h(h_sums.data(), d_sums.p, sizeof(sc_st) * 5 * C, s);
    rt_sync(s);
    std::vector<sc> t0(C), t1(C), t2(C);
    std::vector<sc_st> h_t12(2 * (size_t)C);
    for (int c = 0; c < C; c++) {
        sc tt; st_to_sc(t0[c], h_tsum[c]); st_to_sc(tt, h_tsum[C + c]); st_to_sc(t2[c], h_tsum[2 * C + c]);
        sc_sub(tt, tt, t0[c]); sc_sub(t1[c], tt, t2[c]);
        sc_to_st(h_t12[c], t1[c]); sc_to_st(h_t12[C + c], t2[c]);
    }
    dev_buf d_t12(sizeof(sc_st) * 2 * C, s), d_T12(64 * (size_t)C, s);
    rt_h2d(d_t12.p, h_t12.data(), sizeof(sc_st) * 2 * C, s);
    {
        finalize_args f = {}; f.sBa = d_t12.as<sc_st>(); f.sHa = d_sums.as<sc_st>() + 2 * C; f.tabB = e.tabB; f.tabH = e.tabH;
        f.out32 = d_T12.as<uint8_t>(); f.count = 2 * C;
        run_finalize(s, f);
    }
    std::vector<uint8_t> hT(64 * (size_t)C);
    rt_d2h(hT.data(), d_T12.p, hT.size(), s);
    rt_sync(s);
    // ---- x, shares, w
    std::vector<sc> x(C), w(C);
    std::vector<sc_st> h_x(C), h_w2(2 * (size_t)C);
    serial_for(C, [&](size_t c) {
        transcript &t = ts[c]; uint8_t *o = h_proofs + plen * c;
        memcpy(o + 64, &hT[32 * c], 32); memcpy(o + 96, &hT[32 * (C + c)], 32);
        transcript_append(t, "T_1", o + 64, 32); transcript_append(t, "T_2", o + 96, 32);
        ts_challenge_scalar(t, "x", x[c]);
        sc xx, tx, txb, eb, tmp, sa, ss, st1, st2, szg;
        sc_mul(xx, x[c], x[c]);
        st_to_sc(sa, h_sums[c]); st_to_sc(ss, h_sums[C + c]); st_to_sc(st1, h_sums[2 * C + c]); st_to_sc(st2, h_sums[3 * C + c]); st_to_sc(szg, h_sums[4 * C + c]);
        sc_mul(tmp, t1[c], x[c]); sc_add(tx, t0[c], tmp); sc_mul(tmp, t2[c], xx); sc_add(tx, tx, tmp);
        sc_mul(tmp, st1, x[c]); sc_add(txb, szg, tmp); sc_mul(tmp, st2, xx); sc_add(txb, txb, tmp);
        sc_mul(tmp, ss, x[c]); sc_add(eb, sa, tmp);
        sc_tobytes(o + 128, tx); sc_tobytes(o + 160, txb); sc_tobytes(o + 192, eb);
        transcript_append(t, "t_x", o + 128, 32); transcript_append(t, "t_x_blinding", o + 160, 32); transcript_append(t, "e_blinding", o + 192, 32);
        ts_challenge_scalar(t, "w", w[c]);
        transcript_append(t, "dom-sep", (const uint8_t *)"ipp v1", 6); transcript_append_u64(t, "n", (uint64_t)N);
        sc_to_st(h_x[c], x[c]); sc_to_st(h_w2[c], w[c]); sc_to_st(h_w2[C + c], w[c]);
    });
    dev_buf d_x(sizeof(sc_st) * C, s), d_w2(sizeof(sc_st) * 2 * C, s), d_yinv(sizeof(sc_st) * NT, s);
    rt_h2d(d_x.p, h_x.data(), sizeof(sc_st) * C, s); rt_h2d(d_w2.p, h_w2.data(), sizeof(sc_st) * 2 * C, s);
    LAUNCH(k_lr, dim3((unsigned)((NT + 255) / 256)), dim3(256), s, d_a.as<sc_st>(), d_b.as<sc_st>(), d_sLR.as<sc_st>(), d_yinv.as<sc_st>(), d_x.as<sc_st>(), yitab, N, NT);
    tr.mark("pre_ipp");
    // ---- inner product argument
    // folded generators [C][half]; a tail that starts at round 0 keeps all N original generators there instead
    const size_t half = (N / 2 <= (size_t)std::min(e.tail_np, TAIL_MAX_F / 2)) ? N : (N / 2 ? N / 2 : 1);
    dev_buf d_Gf(sizeof(p3_st) * half * C, s), d_Hf(sizeof(p3_st) * half * C, s);
    sc_st *msmL = d_sLR.as<sc_st>(), *msmR = d_sLR.as<sc_st>() + NT;         // s_L / s_R are dead after k_lr: reuse as MSM scalar buffers
    dev_buf d_cLR(sizeof(sc_st) * 2 * C, s), d_LR(64 * (size_t)C, s), d_u2(sizeof(sc_st) * C, s), d_uinv2(sizeof(sc_st) * C, s), d_nafs(512 * (size_t)C, s);
    std::vector<sc> uprod(C), uinvprod(C);
    for (int c = 0; c < C; c++) { sc_from_u64(uprod[c], 1); sc_from_u64(uinvprod[c], 1); }
    pinned_buf pinLR(e, 64 * (size_t)C); uint8_t *hLR = pinLR.as<uint8_t>();
    std::vector<sc> u(C), uinv(C);
    std::vector<sc_st> h_u2(C), h_uinv2(C); std::vector<int8_t> h_nafs(512 * (size_t)C);
    // RT path: the first r_unf rounds take L/R as table MSMs over the ORIGINAL generators (no generator folding), then one
    // catch-up fold builds G"/H" of length N >> r_unf directly from the tables
    // rounds with half-size <= tail_np run in the fused on-device tail kernel; `pre` rounds come before it
    const size_t tail_np = (size_t)std::min(e.tail_np, TAIL_MAX_F / 2);
    int pre = 0; while (pre < lgN && ((N / 2) >> pre) > tail_np) pre++;
    const int r_unf = rt ? std::min({e.rt_unfold, lgN, pre}) : 0;
    const int nbU = rt ? rt_blocks(N, 2 * C) : 1;
    const uint32_t cstride = 1u << (r_unf > 0 ? r_unf : 0);
    std::vector<sc> cG((size_t)C * cstride), cH((size_t)C * cstride);
    std::vector<sc_st> h_cGH(2 * (size_t)C * cstride);
    for (int c = 0; c < C; c++) { sc_from_u64(cG[(size_t)c * cstride], 1); sc_from_u64(cH[(size_t)c * cstride], 1); }
    dev_buf d_cGH(sizeof(sc_st) * 2 * (size_t)C * cstride, s), d_partU(sizeof(p3_st) * 2 * (size_t)C * nbU, s), d_digs(sizeof(int16_t) * 2 * (size_t)C * cstride * RT_MAXW, s);
    const int nbQ = (int)std::min<size_t>(256, (N / 2 + 255) / 256);
    dev_buf d_partQ(sizeof(sc_st) * 2 * (size_t)C * nbQ, s);
    // frozen level (kernels.cuh K6c): rounds ra .. pre-1 run over the generators as they are at round ra (FA of G" and of H" per chunk)
    int ra = -1; size_t FA = 0; uint32_t cAstride = 1;
    if (e.use_frz) {
        int r = std::max(r_unf, 1);
        while (r < pre && 2 * ((N / 2) >> r) > FRZ_MAX_F) r++;
        const size_t fa = r < lgN ? 2 * ((N / 2) >> r) : 0;
        const double need = (double)C * 2 * fa * FRZ_Q * (FRZ_E + 1) * sizeof(p3_st);
        // (the budget check uses the free memory recorded when the tables were built: cudaMemGetInfo itself blocks for tens of
        //  milliseconds every now and then -- it was the cause of the "slow steps" of DESIGN.md section 7)
        if (e.free_hint.load() == 0) e.free_hint = rt_free_mem();
        if (pre - r >= 2 && fa >= 4 && need < 0.25 * (double)e.free_hint.load()) { ra = r; FA = fa; cAstride = (uint32_t)(FA / ((N / 2) >> (pre - 1))); }
    }
    std::vector<sc> cAG((size_t)C * cAstride), cAH((size_t)C * cAstride);
    std::vector<sc_st> h_cA(2 * (size_t)C * cAstride);
    for (int c = 0; c < C; c++) { sc_from_u64(cAG[(size_t)c * cAstride], 1); sc_from_u64(cAH[(size_t)c * cAstride], 1); }
    dev_buf d_cA(sizeof(sc_st) * 2 * (size_t)C * cAstride, s), d_frzV(sizeof(p3_st) * 2 * (size_t)C * 8, s);
    dev_buf d_frzT(ra >= 0 ? sizeof(p3_st) * (size_t)C * 2 * FA * FRZ_Q * FRZ_E : 16, s);
    int round = 0;
    bool tail_done = false;
    for (size_t np = N / 2; np >= 1; np /= 2, round++) {
        if (round == r_unf) tr.mark("unfolded");
        if (round == ra) {           // enter the frozen level: Straus tables of the current G", H"
            dev_buf d_bases(sizeof(p3_st) * (size_t)C * 2 * FA * FRZ_Q, s);      // returned stream-ordered: the kernels below may still be running
            void *tk = rt_prof_begin(PROF_FRZ, s);
            LAUNCH(k_frz_bases, dim3((unsigned)((2 * FA + 127) / 128), C), dim3(128), s, d_bases.as<p3_st>(), d_Gf.as<p3_st>(), d_Hf.as<p3_st>(), (uint32_t)FA, (uint32_t)half);
            const size_t cnt = (size_t)C * 2 * FA * FRZ_Q;
            LAUNCH(k_frz_tables, dim3((unsigned)((cnt + 127) / 128)), dim3(128), s, d_frzT.as<p3_st>(), d_bases.as<p3_st>(), cnt);
            rt_prof_end(PROF_FRZ, tk, s);
        }
        const bool frozen = ra >= 0 && round >= ra;
        if (round == pre) tr.mark("middle");
        if (round == pre && frozen) {       // leave the frozen level: the 2*np generators the tail starts from, straight from the tables
            const uint32_t nblk = (uint32_t)(FA / (2 * np));
            std::vector<int8_t> h_dg(2 * (size_t)C * nblk * 64);
            for (int c = 0; c < C; c++) for (uint32_t t = 0; t < nblk; t++) {
                sc_radix16(&h_dg[(((size_t)c * 2 + 0) * nblk + t) * 64], cAG[(size_t)c * cAstride + t]);
                sc_radix16(&h_dg[(((size_t)c * 2 + 1) * nblk + t) * 64], cAH[(size_t)c * cAstride + t]);
            }
            dev_buf d_dg(h_dg.size(), s); rt_h2d(d_dg.p, h_dg.data(), h_dg.size(), s);
            frz_exit_args xa = {}; xa.T = d_frzT.as<p3_st>(); xa.digs = d_dg.as<int8_t>(); xa.Gf = d_Gf.as<p3_st>(); xa.Hf = d_Hf.as<p3_st>();
            xa.F = (uint32_t)FA; xa.Fo = (uint32_t)(2 * np); xa.nblk = nblk; xa.stride = (uint32_t)half;
            dev_buf d_xV(sizeof(p3_st) * (size_t)C * 2 * xa.Fo * 8, s); xa.V = d_xV.as<p3_st>();
            void *tk = rt_prof_begin(PROF_FRZ, s);
            LAUNCH_COOP(k_frz_exit, dim3((unsigned)(2 * np), C, 2), dim3(128), s, xa);
            LAUNCH(k_frz_exit_chain, dim3((unsigned)(((size_t)C * 2 * xa.Fo + 127) / 128)), dim3(128), s, xa, (uint32_t)C);
            rt_prof_end(PROF_FRZ, tk, s);
        }
        if (round == pre) tr.mark("exit");
        if (round == pre) {          // np <= tail_np: every remaining round in one launch (kernels.cuh, k_ipp_tail)
            const int rounds_left = lgN - round; const uint32_t ostride = 64 * (uint32_t)rounds_left + 64;
            std::vector<sc_st> h_up(2 * (size_t)C);
            for (int c = 0; c < C; c++) { sc_to_st(h_up[c], uprod[c]); sc_to_st(h_up[C + c], uinvprod[c]); }
            dev_buf d_ts(sizeof(transcript) * (size_t)C, s), d_up(sizeof(sc_st) * 2 * (size_t)C, s), d_tail((size_t)ostride * C, s);
            rt_h2d(d_ts.p, ts.data(), sizeof(transcript) * (size_t)C, s); rt_h2d(d_up.p, h_up.data(), sizeof(sc_st) * 2 * (size_t)C, s);
            // freeze the 2*np generators the tail works on: Straus tables (kernels.cuh K6c), built once for all its rounds
            const uint32_t Ft = (uint32_t)(2 * np);
            if (round == 0) LAUNCH(k_niels_to_p3, dim3((Ft + 127) / 128, C), dim3(128), s, d_Gf.as<p3_st>(), d_Hf.as<p3_st>(), g.G, g.H, Ft, (uint32_t)half);
            dev_buf d_tb(sizeof(p3_st) * (size_t)C * 2 * Ft * FRZ_Q, s), d_tT(sizeof(p3_st) * (size_t)C * 2 * Ft * FRZ_Q * FRZ_E, s);
            void *tk = rt_prof_begin(PROF_TAIL, s);
            LAUNCH(k_frz_bases, dim3((2 * Ft + 127) / 128, C), dim3(128), s, d_tb.as<p3_st>(), d_Gf.as<p3_st>(), d_Hf.as<p3_st>(), Ft, (uint32_t)half);
            LAUNCH(k_frz_tables, dim3((unsigned)(((size_t)C * 2 * Ft * FRZ_Q + 127) / 128)), dim3(128), s, d_tT.as<p3_st>(), d_tb.as<p3_st>(), (size_t)C * 2 * Ft * FRZ_Q);
            tail_args ta = {}; ta.T = d_tT.as<p3_st>();
            ta.a = d_a.as<sc_st>(); ta.b = d_b.as<sc_st>(); ta.yinv = d_yinv.as<sc_st>(); ta.N = N;
            ta.ts = d_ts.as<transcript>(); ta.w = d_w2.as<sc_st>(); ta.uprod = d_up.as<sc_st>(); ta.uinvprod = d_up.as<sc_st>() + C; ta.tabB = e.tabB;
            dev_buf d_tscr(sizeof(p3_st) * 512 * (size_t)C, s); ta.scratch = d_tscr.as<p3_st>();
            ta.out = d_tail.as<uint8_t>(); ta.out_stride = ostride; ta.F = Ft;
            LAUNCH_COOP(k_ipp_tail, dim3(C), dim3(TAIL_THREADS), s, ta);
            rt_prof_end(PROF_TAIL, tk, s);
            std::vector<uint8_t> h_tail((size_t)ostride * C);
            rt_d2h(h_tail.data(), d_tail.p, h_tail.size(), s); rt_sync(s);
            for (int c = 0; c < C; c++) memcpy(h_proofs + plen * c + 224 + 64 * (size_t)round, &h_tail[(size_t)ostride * c], ostride);
            tail_done = true; tr.mark("tail");
            break;
        }
        const int nbI = (int)std::min<size_t>(256, (np + 255) / 256);
        const bool unfolded = round < r_unf;
        if (frozen) {
            const uint32_t nblk = (uint32_t)(FA / (2 * np)); const int nbQA = (int)std::min<size_t>(256, (FA / 2 + 255) / 256);
            for (int c = 0; c < C; c++) for (uint32_t t = 0; t < nblk; t++) { sc_to_st(h_cA[(size_t)c * cAstride + t], cAG[(size_t)c * cAstride + t]); sc_to_st(h_cA[((size_t)C + c) * cAstride + t], cAH[(size_t)c * cAstride + t]); }
            rt_h2d(d_cA.p, h_cA.data(), sizeof(sc_st) * h_cA.size(), s);
            LAUNCH_COOP(k_ipp_scalars_unf, dim3(nbQA, C), dim3(256), s, d_a.as<sc_st>(), d_b.as<sc_st>(), d_yinv.as<sc_st>(), d_cA.as<sc_st>(), d_cA.as<sc_st>() + (size_t)C * cAstride, cAstride,
                        msmL, msmR, d_partQ.as<sc_st>(), N, (uint32_t)np, (uint32_t)FA);
            LAUNCH_COOP(k_sc_sum, dim3(C), dim3(256), s, d_cLR.as<sc_st>(), d_partQ.as<sc_st>(), nbQA, 2);
            frz_reduce_args ra_ = {}; ra_.T = d_frzT.as<p3_st>(); ra_.msmL = msmL; ra_.msmR = msmR; ra_.V = d_frzV.as<p3_st>(); ra_.F = (uint32_t)FA; ra_.np = (uint32_t)np; ra_.C = (uint32_t)C;
            void *tk = rt_prof_begin(PROF_FRZ, s);
            LAUNCH_COOP(k_frz_reduce, dim3(8, 2 * C), dim3(128), s, ra_);
            rt_prof_end(PROF_FRZ, tk, s);
            finalize_args f = {}; f.windows = d_frzV.as<p3_st>(); f.c = 4; f.nw = 8; f.slices = 1; f.sBa = d_cLR.as<sc_st>(); f.sBb = d_w2.as<sc_st>(); f.tabB = e.tabB; f.tabH = e.tabH;
            f.out32 = d_LR.as<uint8_t>(); f.count = 2 * C;
            run_finalize(s, f);
        } else if (unfolded) {
            const uint32_t nblk = 1u << round;
            for (int c = 0; c < C; c++) for (uint32_t t = 0; t < nblk; t++) { sc_to_st(h_cGH[(size_t)c * cstride + t], cG[(size_t)c * cstride + t]); sc_to_st(h_cGH[((size_t)C + c) * cstride + t], cH[(size_t)c * cstride + t]); }
            rt_h2d(d_cGH.p, h_cGH.data(), sizeof(sc_st) * h_cGH.size(), s);
            LAUNCH_COOP(k_ipp_scalars_unf, dim3(nbQ, C), dim3(256), s, d_a.as<sc_st>(), d_b.as<sc_st>(), d_yinv.as<sc_st>(), d_cGH.as<sc_st>(), d_cGH.as<sc_st>() + (size_t)C * cstride, cstride,
                        msmL, msmR, d_partQ.as<sc_st>(), N, (uint32_t)np, (uint32_t)N);
            LAUNCH_COOP(k_sc_sum, dim3(C), dim3(256), s, d_cLR.as<sc_st>(), d_partQ.as<sc_st>(), nbQ, 2);
            rt_msm_args aL = {}; aL.scalars = msmL; aL.T = (uint32_t)N; aL.scalar_stride = (uint32_t)N; aL.nG = (uint32_t)(N / 2); aL.np = (uint32_t)np; aL.mode = 1; aL.rt = *rt; aL.partial = d_partU.as<p3_st>();
            rt_msm_args aR = aL; aR.scalars = msmR; aR.mode = 2; aR.partial = d_partU.as<p3_st>() + (size_t)C * nbU;
            run_rt_msm(e, s, aL, nbU, C); run_rt_msm(e, s, aR, nbU, C);
            finalize_args f = {}; f.partial = d_partU.as<p3_st>(); f.npartial = nbU; f.sBa = d_cLR.as<sc_st>(); f.sBb = d_w2.as<sc_st>(); f.tabB = e.tabB; f.tabH = e.tabH;
            f.out32 = d_LR.as<uint8_t>(); f.count = 2 * C;
            run_finalize(s, f);
        } else {
            LAUNCH_COOP(k_ipp_scalars, dim3(nbI, C), dim3(256), s, d_a.as<sc_st>(), d_b.as<sc_st>(), d_yinv.as<sc_st>(), msmL, msmR, d_part.as<sc_st>(), N, (uint32_t)np);
            LAUNCH_COOP(k_sc_sum, dim3(C), dim3(256), s, d_cLR.as<sc_st>(), d_part.as<sc_st>(), nbI, 2);
            // L and R of every chunk in one launch: msm = lr*C + c
            const msm_plan pl = msm_plan_for(2 * np, 2 * (size_t)C);
            dev_buf d_win(sizeof(p3_st) * pl.out_count(2 * (size_t)C), s);
            msm_args a = {}; a.v[0].scalars = msmL; a.v[1].scalars = msmR; a.split = (uint32_t)C;
            a.T = (uint32_t)(2 * np); a.scalar_stride = (uint32_t)(2 * np); a.nseg = 2; a.out = d_win.as<p3_st>();
            if (round == 0) {
                a.v[0].seg[0] = mk_seg(g.G + np, (uint32_t)np, 0, 0); a.v[0].seg[1] = mk_seg(g.H, (uint32_t)np, 0, 0);
                a.v[1].seg[0] = mk_seg(g.G, (uint32_t)np, 0, 0);      a.v[1].seg[1] = mk_seg(g.H + np, (uint32_t)np, 0, 0);
            } else {
                a.v[0].seg[0] = mk_seg(d_Gf.as<p3_st>() + np, (uint32_t)np, (uint32_t)half, 1); a.v[0].seg[1] = mk_seg(d_Hf.as<p3_st>(), (uint32_t)np, (uint32_t)half, 1);
                a.v[1].seg[0] = mk_seg(d_Gf.as<p3_st>(), (uint32_t)np, (uint32_t)half, 1);      a.v[1].seg[1] = mk_seg(d_Hf.as<p3_st>() + np, (uint32_t)np, (uint32_t)half, 1);
            }
            run_msm(e, s, a, pl, 2 * (uint32_t)C);
            finalize_args f = {}; fin_windows(f, d_win.as<p3_st>(), pl); f.sBa = d_cLR.as<sc_st>(); f.sBb = d_w2.as<sc_st>(); f.tabB = e.tabB; f.tabH = e.tabH;
            f.out32 = d_LR.as<uint8_t>(); f.count = 2 * C;
            run_finalize(s, f);
        }
        rt_d2h(hLR, d_LR.p, 64 * (size_t)C, s);
        rt_sync(s);
        serial_for(C, [&](size_t c) {
            transcript &t = ts[c]; uint8_t *o = h_proofs + plen * c + 224 + 64 * round;
            memcpy(o, &hLR[32 * c], 32); memcpy(o + 32, &hLR[32 * (C + c)], 32);
            transcript_append(t, "L", o, 32); transcript_append(t, "R", o + 32, 32);
            ts_challenge_scalar(t, "u", u[c]);
        });
        uinv = u; sc_batch_invert(uinv);
        const int lgnp = ilog2_sz(np);
        for (int c = 0; c < C; c++) {
            sc u2, ui2, sH, yp;
            sc_mul(u2, u[c], u[c]); sc_mul(ui2, uinv[c], uinv[c]);
            st_to_sc(yp, h_yinvpow2[32 * c + lgnp]); sc_mul(sH, ui2, yp);            // u^-2 y^-np
            sc_to_st(h_u2[c], u2); sc_to_st(h_uinv2[c], ui2);
            if (frozen) {
                const uint32_t nblk = (uint32_t)(FA / (2 * np)); sc *g0 = &cAG[(size_t)c * cAstride], *h0 = &cAH[(size_t)c * cAstride];
                if (2 * nblk <= cAstride) for (uint32_t t = nblk; t-- > 0;) { sc gt = g0[t], ht = h0[t]; g0[2 * t] = gt; sc_mul(g0[2 * t + 1], gt, u2); h0[2 * t] = ht; sc_mul(h0[2 * t + 1], ht, sH); }
            } else if (unfolded) {       // coefficient tables of the next level: c'[2t] = c[t], c'[2t+1] = c[t] * s
                const uint32_t nblk = 1u << round; sc *g0 = &cG[(size_t)c * cstride], *h0 = &cH[(size_t)c * cstride];
                for (uint32_t t = nblk; t-- > 0;) { sc gt = g0[t], ht = h0[t]; g0[2 * t] = gt; sc_mul(g0[2 * t + 1], gt, u2); h0[2 * t] = ht; sc_mul(h0[2 * t + 1], ht, sH); }
            } else { sc_naf(&h_nafs[512 * c], u2, FOLD_W); sc_naf(&h_nafs[512 * c + 256], sH, FOLD_W); }
            sc_mul(uprod[c], uprod[c], u[c]); sc_mul(uinvprod[c], uinvprod[c], uinv[c]);
        }
        rt_h2d(d_u2.p, h_u2.data(), sizeof(sc_st) * C, s); rt_h2d(d_uinv2.p, h_uinv2.data(), sizeof(sc_st) * C, s);
        LAUNCH(k_ipp_fold_scalars, dim3((unsigned)((np + 255) / 256), C), dim3(256), s, d_a.as<sc_st>(), d_b.as<sc_st>(), d_u2.as<sc_st>(), d_uinv2.as<sc_st>(), N, (uint32_t)np);
        if (frozen) {
            // nothing to fold: the generators stay frozen, only the coefficient tables grew
        } else if (unfolded) {
            if (round + 1 == r_unf && np >= 2) {          // catch-up: G", H" of length np straight from the tables
                const uint32_t nblk = 1u << r_unf;
                std::vector<int16_t> h_digs(2 * (size_t)C * nblk * RT_MAXW);
                for (int c = 0; c < C; c++) for (uint32_t t = 0; t < nblk; t++) {
                    rt_digits(&h_digs[(((size_t)c * 2 + 0) * nblk + t) * RT_MAXW], cG[(size_t)c * cstride + t], *rt);
                    rt_digits(&h_digs[(((size_t)c * 2 + 1) * nblk + t) * RT_MAXW], cH[(size_t)c * cstride + t], *rt);
                }
                rt_h2d(d_digs.p, h_digs.data(), sizeof(int16_t) * h_digs.size(), s);
                catchup_args ca = {}; ca.rt = *rt; ca.Gf = d_Gf.as<p3_st>(); ca.Hf = d_Hf.as<p3_st>(); ca.digits = d_digs.as<int16_t>(); ca.nr = (uint32_t)np; ca.nblk = nblk; ca.stride = (uint32_t)half;
                void *tk = rt_prof_begin(PROF_FOLD, s);
                LAUNCH_COOP(k_rt_catchup, dim3((unsigned)((np + 127) / 128), C, 2), dim3(128), s, ca);
                rt_prof_end(PROF_FOLD, tk, s);
            }
        } else if (np >= 2) {
            rt_h2d(d_nafs.p, h_nafs.data(), h_nafs.size(), s);
            fold_args fa = {}; fa.Gn = round == 0 ? g.G : nullptr; fa.Hn = round == 0 ? g.H : nullptr;
            fa.Gf = d_Gf.as<p3_st>(); fa.Hf = d_Hf.as<p3_st>(); fa.nafs = d_nafs.as<int8_t>(); fa.np = (uint32_t)np; fa.stride = (uint32_t)half;
            void *tk = rt_prof_begin(PROF_FOLD, s);
            LAUNCH_COOP(k_ipp_fold_points, dim3((unsigned)((np + 127) / 128), C, 2), dim3(128), s, fa);
            rt_prof_end(PROF_FOLD, tk, s);
        }
    }
    if (tail_done) return;
    // ---- final a, b: a = a^ prod u_k, b = b^ prod u_k^-1
    std::vector<sc_st> h_ab(2 * (size_t)C);
    for (int c = 0; c < C; c++) { rt_d2h(&h_ab[2 * c], d_a.as<sc_st>() + (size_t)c * N, sizeof(sc_st), s); rt_d2h(&h_ab[2 * c + 1], d_b.as<sc_st>() + (size_t)c * N, sizeof(sc_st), s); }
    rt_sync(s);
    for (int c = 0; c < C; c++) {
        sc a, b; st_to_sc(a, h_ab[2 * c]); st_to_sc(b, h_ab[2 * c + 1]);
        sc_mul(a, a, uprod[c]); sc_mul(b, b, uinvprod[c]);
        uint8_t *o = h_proofs + plen * c + 224 + 64 * lgN;
        sc_tobytes(o, a); sc_tobytes(o + 32, b);
    }
}


// run f(group, c0, c1, stream) for `groups` contiguous chunk ranges concurrently (one host thread + one stream per group)
template <class F> static void for_chunk_groups(rofl_engine &e, size_t C, F f) {
    size_t G = std::max<size_t>(1, std::min<size_t>({(size_t)e.groups, e.gstreams.size(), C}));
    if (G == 1) { f(0, (size_t)0, C, e.stream); return; }
    std::vector<std::thread> th; std::vector<std::string> errs(G);
    for (size_t gi = 0; gi < G; gi++) th.emplace_back([&, gi] {
        try { rt_set_device(e.device); f(gi, C * gi / G, C * (gi + 1) / G, e.gstreams[gi]); } catch (const std::exception &ex) { errs[gi] = ex.what(); }
    });
    for (auto &t : th) t.join();
    for (auto &m : errs) if (!m.empty()) throw std::runtime_error(m);
}

// =============================================================================================================================
// range_proof_vec::create_rangeproof (range_proof_vec/mod.rs:16-102).  d_values / d_blind / d_commits are device pointers.
// returns 0 ok, 2 ValueOutOfRangeError, -1 InvalidBitsize, -2 bad arguments, -98 NaN input (reference panics),
// -99 non power-of-two chunking (reference panics "Should not get here")
// =============================================================================================================================
// A shard = chunks [chunk_begin, chunk_begin + n_chunks) of a larger update (chunk length m): the caller passes only that
// slice of the values / blindings (D = its real elements, the rest of the m * n_chunks positions is the reference's zero
// padding) and gets exactly the proofs the whole-update call would produce for those chunks (same per-chunk nonce streams).
struct shard_spec { size_t m, chunk_begin, n_chunks; };
static int engine_range_prove(rofl_engine &e, const float *d_values, const uint8_t *d_blind, size_t D, int range, size_t n_partition,
                              int n_bits, int frac, const uint8_t seed[32], uint8_t *h_proofs, size_t *proof_len, size_t *n_proofs, uint8_t *d_commits,
                              const shard_spec *shard = nullptr) {
    if (!fp_ok(n_bits, frac) || range < 1 || range > n_bits) return -2;
    if (shard ? (shard->m == 0 || shard->n_chunks == 0 || D > shard->m * shard->n_chunks) : (D == 0 || n_partition == 0)) return -2;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    const size_t Dp = shard ? shard->m * shard->n_chunks : next_pow2_sz(D);
    const size_t C = shard ? shard->n_chunks : std::min(Dp, n_partition), m = Dp / C;                         // :54-55
    const size_t c_off = shard ? shard->chunk_begin : 0;
    const bool bitsize_ok = (range == 8 || range == 16 || range == 32 || range == 64);
    const bool chunk_ok = !(m & (m - 1)) && m * C == Dp;
    dev_buf d_V(32 * Dp, s), d_vals(8 * Dp, s), d_bl(sizeof(sc_st) * Dp, s), d_flags(sizeof(int), s);
    rt_memset(d_flags.p, 0, sizeof(int), s);
    commit_args ca = {}; ca.values = d_values; ca.blind = d_blind; ca.D = D; ca.Dp = Dp; ca.n_bits = n_bits; ca.frac = frac;
    ca.shift_bits = range; ca.mx = clip_max_f(range, n_bits, frac); ca.mn = -ca.mx; ca.tabB = e.tabB; ca.tabH = e.tabH;
    ca.V = d_V.as<uint8_t>(); ca.C = d_commits; ca.vals = d_vals.as<uint64_t>(); ca.blind_sc = d_bl.as<sc_st>(); ca.flags = d_flags.as<int>();
    void *tk = rt_prof_begin(PROF_COMMIT, s);
    LAUNCH(k_commit, dim3((unsigned)((Dp + 127) / 128)), dim3(128), s, ca);
    rt_prof_end(PROF_COMMIT, tk, s);
    int flags = 0; rt_d2h(&flags, d_flags.p, sizeof(int), s); rt_sync(s);
    if (flags & 2) return 2;                                                         // :26-29
    if (flags & 1) return -98;
    if (!bitsize_ok) return -1;
    if (!chunk_ok) return -99;
    gens_entry &g = engine_gens(e, range, (int)m);
    rt_tables rt; const bool have_rt = engine_rt(e, g, range, (int)m, rt);
    std::vector<uint8_t> keys(32 * C);
    for (size_t c = 0; c < C; c++) derive_key(&keys[32 * c], seed, DOM_RANGE_PROVE, c_off + c);
    const size_t plen = 32 * (9 + 2 * (size_t)ilog2_sz((size_t)range * m));
    for_chunk_groups(e, C, [&](size_t, size_t c0, size_t c1, cudaStream_t gs) {
        std::vector<uint8_t> k(keys.begin() + 32 * c0, keys.begin() + 32 * c1);
        prove_chunks(e, gs, "RangeProof", range, (int)m, (int)(c1 - c0), g, have_rt ? &rt : nullptr, d_vals.as<uint64_t>() + c0 * m, d_bl.as<sc_st>() + c0 * m, d_V.as<uint8_t>() + 32 * c0 * m, k, h_proofs + plen * c0);
    });
    *proof_len = plen; *n_proofs = C;
    return 0;
}

// =============================================================================================================================
// RangeProof::verify_multiple for C chunks: d_Vp3 / h_V32 hold C*m (shifted) commitments.
// verdict[c] = 1 accept / 0 VerificationError.  returns 0 or a negative error (-1 FormatError).
// =============================================================================================================================
static int verify_chunks(rofl_engine &e, cudaStream_t s, const char *label, int n, int m, int C, const gens_entry &g, const rt_tables *rt, const p3_st *d_Vp3, const uint8_t *h_V32,
                         const uint8_t *h_proofs, size_t plen, const std::vector<uint8_t> &keys, std::vector<int> &verdict) {
    verdict.assign(C, 0);
    phase_trace tr(s);
    // RangeProof::from_bytes / InnerProductProof::from_bytes
    if (plen % 32 || plen < 7 * 32) return -1;
    size_t ne = plen / 32 - 7;
    if (ne < 2 || (ne - 2) % 2) return -1;
    const size_t lg = (ne - 2) / 2; if (lg >= 32) return -1;
    for (int c = 0; c < C; c++) {
        const uint8_t *p = h_proofs + plen * c; sc t;
        for (size_t off : {(size_t)128, (size_t)160, (size_t)192, 224 + 64 * lg, 224 + 64 * lg + 32}) { sc_frombytes(t, p + off); if (!sc_is_canonical(t)) return -1; }
    }
    const size_t N = (size_t)n * m;
    if ((m & (m - 1)) || N != ((size_t)1 << lg)) return 0;                      // verification_scalars: n != 1 << lg_n -> VerificationError (all chunks false)
    const int lgN = (int)lg, nsmall = 6 + 2 * lgN;
    // all C chunks are checked with ONE equation: sum_c rho_c * (chunk c's mega-check) == identity (kernels.cuh, k_verify_scalars)
    const int chs = 5 + 2 * lgN;
    const uint32_t vstride = (uint32_t)(m + nsmall);
    std::vector<sc_st> h_chal((size_t)C * chs), h_small((size_t)C * nsmall), h_yinvpow2(32 * (size_t)C), h_zpow2(32 * (size_t)C);
    std::vector<uint8_t> h_smallpts(32 * (size_t)C * nsmall);
    std::vector<int> host_ok(C, 1);
    std::vector<sc> y(C), yinv(C), z(C), x(C), w(C), cc(C), rho(C);
    std::vector<std::vector<sc>> u(C, std::vector<sc>(lgN));
    std::vector<sc> allu; allu.reserve((size_t)C * (lgN + 1));
    parallel_for(C, e.host_threads, [&](size_t c) {
        const uint8_t *p = h_proofs + plen * c, *ipp = p + 224;
        transcript t; transcript_init(t, label);
        transcript_append(t, "dom-sep", (const uint8_t *)"rangeproof v1", 13);
        transcript_append_u64(t, "n", (uint64_t)n); transcript_append_u64(t, "m", (uint64_t)m);
        for (int j = 0; j < m; j++) transcript_append(t, "V", h_V32 + 32 * (c * m + j), 32);
        int ok = 1;
        ok &= !is_zero32(p) && !is_zero32(p + 32) && !is_zero32(p + 64) && !is_zero32(p + 96);          // validate_and_append_point
        transcript_append(t, "A", p, 32); transcript_append(t, "S", p + 32, 32);
        ts_challenge_scalar(t, "y", y[c]); ts_challenge_scalar(t, "z", z[c]);
        transcript_append(t, "T_1", p + 64, 32); transcript_append(t, "T_2", p + 96, 32);
        ts_challenge_scalar(t, "x", x[c]);
        transcript_append(t, "t_x", p + 128, 32); transcript_append(t, "t_x_blinding", p + 160, 32); transcript_append(t, "e_blinding", p + 192, 32);
        ts_challenge_scalar(t, "w", w[c]);
        uint32_t kw[8]; key_words(kw, &keys[32 * c]); nonce_scalar(cc[c], kw, 0);                     // batching scalar c <- rng
        nonce_scalar(rho[c], kw, 1);                                                                   // cross-chunk weight
        transcript_append(t, "dom-sep", (const uint8_t *)"ipp v1", 6); transcript_append_u64(t, "n", (uint64_t)N);
        for (int k = 0; k < lgN; k++) {
            ok &= !is_zero32(ipp + 64 * k) && !is_zero32(ipp + 64 * k + 32);
            transcript_append(t, "L", ipp + 64 * k, 32); transcript_append(t, "R", ipp + 64 * k + 32, 32);
            ts_challenge_scalar(t, "u", u[c][k]);
        }
        host_ok[c] = ok;
        // small points: A S T1 T2 L.. R.. H B
        uint8_t *sp = &h_smallpts[32 * c * nsmall];
        memcpy(sp, p, 128);
        for (int k = 0; k < lgN; k++) { memcpy(sp + 32 * (4 + k), ipp + 64 * k, 32); memcpy(sp + 32 * (4 + lgN + k), ipp + 64 * k + 32, 32); }
        memcpy(sp + 32 * (4 + 2 * lgN), e.H32, 32); memcpy(sp + 32 * (5 + 2 * lgN), e.B32, 32);
    });
    tr.mark("v_transcripts");
    for (int c = 0; c < C; c++) { allu.push_back(y[c]); for (int k = 0; k < lgN; k++) allu.push_back(u[c][k]); }
    for (auto &v : allu) if (sc_iszero(v)) sc_from_u64(v, 1);                    // (probability 2^-252; keeps the batch inversion defined)
    sc_batch_invert(allu);
    parallel_for(C, e.host_threads, [&](size_t c) {
        const uint8_t *p = h_proofs + plen * c, *ipp = p + 224;
        const sc *inv = &allu[c * (lgN + 1)];
        const sc &r = rho[c];
        yinv[c] = inv[0];
        sc t_x, t_xb, e_bl, a, b, zz, tmp, tmp2;
        sc_frombytes(t_x, p + 128); sc_frombytes(t_xb, p + 160); sc_frombytes(e_bl, p + 192);
        sc_frombytes(a, ipp + 64 * lgN); sc_frombytes(b, ipp + 64 * lgN + 32);
        sc_mul(zz, z[c], z[c]);
        sc_st *ch = &h_chal[c * chs];
        sc_mul(tmp, r, z[c]); sc_to_st(ch[0], tmp); sc_mul(tmp, r, zz); sc_to_st(ch[1], tmp);
        sc_mul(tmp, r, a); sc_to_st(ch[2], tmp); sc_mul(tmp, r, b); sc_to_st(ch[3], tmp); sc_to_st(ch[4], cc[c]);
        sc_st *sm = &h_small[c * nsmall];
        sc cx; sc_mul(cx, cc[c], x[c]);
        auto put = [&](int i, const sc &v) { sc t; sc_mul(t, v, r); sc_to_st(sm[i], t); };                  // every small scalar carries rho_c
        sc_to_st(sm[0], r); put(1, x[c]); put(2, cx); sc_mul(tmp, cx, x[c]); put(3, tmp);
        for (int k = 0; k < lgN; k++) {
            sc_to_st(ch[5 + k], u[c][k]); sc_to_st(ch[5 + lgN + k], inv[1 + k]);
            sc_mul(tmp, u[c][k], u[c][k]); put(4 + k, tmp);
            sc_mul(tmp, inv[1 + k], inv[1 + k]); put(4 + lgN + k, tmp);
        }
        sc_mul(tmp, cc[c], t_xb); sc_add(tmp, tmp, e_bl); sc_neg(tmp, tmp); put(4 + 2 * lgN, tmp);      // H: -e_bl - c t_x_bl
        // delta = (z - zz) sum_{i<N} y^i - z^3 (2^n - 1) sum_{j<m} z^j ; sums of powers via prod (1 + s^(2^b))
        sc_pow2_table(&h_yinvpow2[32 * c], yinv[c]); sc_pow2_table(&h_zpow2[32 * c], z[c]);
        sc one, sum_y, sum_z, pw, delta, s2; sc_from_u64(one, 1); sc_from_u64(sum_y, 1); sc_from_u64(sum_z, 1);
        pw = y[c]; for (int bb = 0; bb < lgN; bb++) { sc_add(tmp, one, pw); sc_mul(sum_y, sum_y, tmp); sc_mul(pw, pw, pw); }
        pw = z[c]; for (int bb = 0; bb < ilog2_sz(m); bb++) { sc_add(tmp, one, pw); sc_mul(sum_z, sum_z, tmp); sc_mul(pw, pw, pw); }
        sc_from_u64(s2, n == 64 ? ~0ULL : ((1ULL << n) - 1));
        sc_sub(delta, z[c], zz); sc_mul(delta, delta, sum_y);
        sc_mul(tmp, zz, z[c]); sc_mul(tmp, tmp, s2); sc_mul(tmp, tmp, sum_z); sc_sub(delta, delta, tmp);
        sc_mul(tmp, a, b); sc_sub(tmp, t_x, tmp); sc_mul(tmp, w[c], tmp);                                        // w (t_x - a b)
        sc_sub(tmp2, delta, t_x); sc_mul(tmp2, cc[c], tmp2); sc_add(tmp, tmp, tmp2); put(5 + 2 * lgN, tmp);      // B
    });
    tr.mark("v_host_scalars");
    const vtab_layout vt = vtab_make(lgN, ilog2_sz(m));
    dev_buf d_chal(sizeof(sc_st) * h_chal.size(), s), d_yinvpow2(sizeof(sc_st) * 32 * C, s), d_zpow2(sizeof(sc_st) * 32 * C, s);
    dev_buf d_tab(sizeof(sc_st) * (size_t)C * vt.total, s), d_gh(sizeof(sc_st) * 2 * N, s), d_var(sizeof(sc_st) * (size_t)C * vstride, s);
    dev_buf d_sp32(h_smallpts.size(), s), d_sp(sizeof(p3_st) * (size_t)C * nsmall, s), d_bad(sizeof(int) * C, s), d_id(sizeof(int), s), d_fix(sizeof(p3_st), s);
    rt_h2d(d_chal.p, h_chal.data(), sizeof(sc_st) * h_chal.size(), s);
    rt_h2d(d_yinvpow2.p, h_yinvpow2.data(), sizeof(sc_st) * 32 * C, s); rt_h2d(d_zpow2.p, h_zpow2.data(), sizeof(sc_st) * 32 * C, s);
    rt_h2d(d_sp32.p, h_smallpts.data(), h_smallpts.size(), s);
    rt_memset(d_bad.p, 0, sizeof(int) * C, s);
    rt_h2d(d_var.as<sc_st>() + (size_t)C * m, h_small.data(), sizeof(sc_st) * h_small.size(), s);
    LAUNCH(k_verify_tables, dim3((vt.total + m + 255) / 256, C), dim3(256), s, d_tab.as<sc_st>(), vt, d_var.as<sc_st>(), (uint32_t)m, d_chal.as<sc_st>(), chs, d_yinvpow2.as<sc_st>(), d_zpow2.as<sc_st>(), m);
    LAUNCH_COOP(k_verify_scalars, dim3((unsigned)((N + 63) / 64)), dim3(256), s, d_gh.as<sc_st>(), d_tab.as<sc_st>(), vt, d_chal.as<sc_st>(), chs, n, C);
    LAUNCH(k_decompress, dim3((unsigned)(((size_t)C * nsmall + 127) / 128)), dim3(128), s, d_sp.as<p3_st>(), (uint8_t *)nullptr, d_sp32.as<uint8_t>(), (size_t)C * nsmall, (size_t)C * nsmall, (const p3_st *)nullptr, d_bad.as<int>(), (size_t)nsmall);
    tr.mark("v_scalar_kernels");
    // fixed generators: one 2N-term MSM (radix-256 tables when they exist, bucket MSM otherwise) -> d_fix
    {
        finalize_args f = {}; f.tabB = e.tabB; f.tabH = e.tabH; f.out_p3 = d_fix.as<p3_st>(); f.count = 1;
        if (rt) {
            const int nbV = rt_blocks(2 * N, 1);
            dev_buf d_partV(sizeof(p3_st) * (size_t)nbV, s);
            rt_msm_args ra = {}; ra.scalars = d_gh.as<sc_st>(); ra.T = (uint32_t)(2 * N); ra.scalar_stride = (uint32_t)(2 * N); ra.nG = (uint32_t)N; ra.mode = 0; ra.rt = *rt; ra.partial = d_partV.as<p3_st>();
            run_rt_msm(e, s, ra, nbV, 1);
            f.partial = d_partV.as<p3_st>(); f.npartial = nbV;
            run_finalize(s, f);
        } else {
            const msm_plan pl = msm_plan_for(2 * N, 1);
            dev_buf d_winF(sizeof(p3_st) * pl.out_count(1), s);
            msm_args a = {}; a.v[0].scalars = d_gh.as<sc_st>(); a.split = 1; a.T = (uint32_t)(2 * N); a.scalar_stride = (uint32_t)(2 * N); a.nseg = 2; a.out = d_winF.as<p3_st>();
            a.v[0].seg[0] = mk_seg(g.G, (uint32_t)N, 0, 0); a.v[0].seg[1] = mk_seg(g.H, (uint32_t)N, 0, 0);
            run_msm(e, s, a, pl, 1);
            fin_windows(f, d_winF.as<p3_st>(), pl);
            run_finalize(s, f);
        }
    }
    tr.mark("v_gen_msm");
    // commitments and proof points of ALL chunks: one sliced MSM over C*m + C*nsmall terms (scalars laid out the same way)
    {
        const uint32_t TV = (uint32_t)((size_t)C * m + (size_t)C * nsmall);
        const msm_plan pl = msm_plan_for(TV, 1);
        dev_buf d_winV(sizeof(p3_st) * pl.out_count(1), s);
        msm_args a = {}; a.v[0].scalars = d_var.as<sc_st>(); a.split = 1; a.T = TV; a.scalar_stride = TV; a.nseg = 2; a.out = d_winV.as<p3_st>();
        a.v[0].seg[0] = mk_seg(d_Vp3, (uint32_t)((size_t)C * m), 0, 1); a.v[0].seg[1] = mk_seg(d_sp.p, (uint32_t)((size_t)C * nsmall), 0, 1);
        run_msm(e, s, a, pl, 1);
        finalize_args f = {}; fin_windows(f, d_winV.as<p3_st>(), pl);
        f.partial = d_fix.as<p3_st>(); f.npartial = 1; f.tabB = e.tabB; f.tabH = e.tabH; f.is_id = d_id.as<int>(); f.count = 1;
        run_finalize(s, f);
    }
    std::vector<int> h_bad(C); int h_id = 0;
    rt_d2h(&h_id, d_id.p, sizeof(int), s); rt_d2h(h_bad.data(), d_bad.p, sizeof(int) * C, s);
    rt_sync(s);
    tr.mark("v_var_msm");
    int all_ok = h_id;
    for (int c = 0; c < C; c++) all_ok &= (host_ok[c] && !h_bad[c]) ? 1 : 0;
    for (int c = 0; c < C; c++) verdict[c] = all_ok;
    return 0;
}

// range_proof_vec::verify_rangeproof (range_proof_vec/mod.rs:149-191).  d_commits: D compressed points (device).
// returns 1 true, 0 false, -1 FormatError, -2 InvalidBitsize / bad args, -3 InvalidGeneratorsLength, -4 undecodable commitment
static int engine_range_verify(rofl_engine &e, const uint8_t *h_proofs, size_t plen, size_t n_proofs, const uint8_t *d_commits, size_t D,
                               int range, const uint8_t seed[32], const shard_spec *shard = nullptr) {
    if (n_proofs == 0 || range < 1 || range > 64) return -2;
    if (shard ? (shard->m == 0 || shard->n_chunks != n_proofs || D > shard->m * shard->n_chunks) : D == 0) return -2;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    const size_t Dp = shard ? shard->m * shard->n_chunks : next_pow2_sz(D), m = Dp / n_proofs;                     // :168
    const size_t c_off = shard ? shard->chunk_begin : 0;
    if (m == 0) return -3;
    // RangeProof::from_bytes happens when the caller deserialises (params.rs:444-458): format errors come first
    if (plen % 32 || plen < 7 * 32) return -1;
    { size_t ne = plen / 32 - 7; if (ne < 2 || (ne - 2) % 2 || (ne - 2) / 2 >= 32) return -1; }
    if (!(range == 8 || range == 16 || range == 32 || range == 64)) {
        // scalars must still be canonical for from_bytes to succeed; check then report InvalidBitsize
        size_t lg = (plen / 32 - 9) / 2;
        for (size_t c = 0; c < n_proofs; c++) { const uint8_t *p = h_proofs + plen * c; sc t; for (size_t off : {(size_t)128, (size_t)160, (size_t)192, 224 + 64 * lg, 224 + 64 * lg + 32}) { sc_frombytes(t, p + off); if (!sc_is_canonical(t)) return -1; } }
        return -2;
    }
    const size_t C = std::min(n_proofs, (Dp + m - 1) / m);                        // zip(proofs, chunks)
    phase_trace tr(s);
    dev_buf d_off(sizeof(p3_st), s), d_offs(sizeof(sc_st), s), d_Vp3(sizeof(p3_st) * Dp, s), d_V32(32 * Dp, s), d_bad(sizeof(int), s);
    { sc o; sc_from_u64(o, 1ULL << (range - 1)); sc_st os; sc_to_st(os, o); rt_h2d(d_offs.p, &os, sizeof(os), s);
      finalize_args f = {}; f.sBa = d_offs.as<sc_st>(); f.tabB = e.tabB; f.tabH = e.tabH; f.out_p3 = d_off.as<p3_st>(); f.count = 1;
      run_finalize(s, f); }
    rt_memset(d_bad.p, 0, sizeof(int), s);
    LAUNCH(k_decompress, dim3((unsigned)((Dp + 127) / 128)), dim3(128), s, d_Vp3.as<p3_st>(), d_V32.as<uint8_t>(), d_commits, D, Dp, d_off.as<p3_st>(), d_bad.as<int>(), Dp);
    std::vector<uint8_t> hV(32 * Dp); int bad = 0;
    rt_d2h(hV.data(), d_V32.p, hV.size(), s); rt_d2h(&bad, d_bad.p, sizeof(int), s);
    rt_sync(s);
    tr.mark("v_decompress");
    if (bad) return -4;
    gens_entry &g = engine_gens(e, range, (int)m);
    rt_tables rt; const bool have_rt = engine_rt(e, g, range, (int)m, rt);
    std::vector<uint8_t> keys(32 * C);
    for (size_t c = 0; c < C; c++) derive_key(&keys[32 * c], seed, DOM_RANGE_VERIFY, c_off + c);
    // one group: every chunk goes into the same batched check (kernels.cuh, k_verify_scalars), splitting would repeat the generator MSM
    std::vector<int> rcs(2, 1);
    [&](auto f) { f((size_t)0, (size_t)0, C, e.stream); }([&](size_t gi, size_t c0, size_t c1, cudaStream_t gs) {
        std::vector<uint8_t> k(keys.begin() + 32 * c0, keys.begin() + 32 * c1); std::vector<int> verdict;
        int rc = verify_chunks(e, gs, "RangeProof", range, (int)m, (int)(c1 - c0), g, have_rt ? &rt : nullptr, d_Vp3.as<p3_st>() + c0 * m, hV.data() + 32 * c0 * m, h_proofs + plen * c0, plen, k, verdict);
        int res = 1; for (int v : verdict) res &= v;                               // :183-190
        rcs[gi] = rc < 0 ? rc : res;
    });
    int res = 1;
    for (int r : rcs) { if (r < 0) return r; res &= r; }
    return res;
}

// =============================================================================================================================
// l2_range_proof_vec::create_rangeproof_l2 (l2_range_proof_vec/mod.rs:15-140): one m=1 proof on sum x_i^2 with blinding sum r_i.
// h_values is a HOST copy of the values: the reference cross-checks the scalar sum against a sequential f32 fold (:44-58)
// whose rounding depends on the summation order, so that O(D) fold runs on the host exactly as written.
// returns 0 ok, 2 ValueOutOfRange, 3 OverflowError, 4 NormOutOfRange, -1 InvalidBitsize, -2 bad args
// =============================================================================================================================
static int engine_l2_prove(rofl_engine &e, const float *h_values, const float *d_values, const uint8_t *d_blind, size_t D, int range,
                           int n_bits, int frac, const uint8_t seed[32], uint8_t *h_proof, size_t *proof_len, uint8_t *h_commit) {
    if (!fp_ok(n_bits, frac) || D == 0 || range < 1) return -2;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    const float mx = clip_max_f(range, n_bits, frac), mn = -mx;
    for (size_t i = 0; i < D; i++) if (mn > h_values[i] || h_values[i] > mx) return 2;
    const int nb = (int)std::min<size_t>(256, (D + 255) / 256);
    dev_buf d_part(sizeof(sc_st) * 2 * nb, s), d_sum(sizeof(sc_st) * 2, s), d_flags(sizeof(int), s);
    rt_memset(d_flags.p, 0, sizeof(int), s);
    LAUNCH_COOP(k_l2_sums, dim3(nb), dim3(256), s, d_part.as<sc_st>(), d_values, d_blind, D, n_bits, frac, d_flags.as<int>());
    LAUNCH_COOP(k_sc_sum, dim3(1), dim3(256), s, d_sum.as<sc_st>(), d_part.as<sc_st>(), nb, 2);
    sc_st hs[2]; int flags = 0;
    rt_d2h(hs, d_sum.p, sizeof(hs), s); rt_d2h(&flags, d_flags.p, sizeof(int), s); rt_sync(s);
    if (flags & 1) return -98;
    sc val, bsum; st_to_sc(val, hs[0]); st_to_sc(bsum, hs[1]);
    // :44-58 sequential f32 cross-check
    const float shift = (float)(1 << frac); float val_float = 0.0f;
    for (size_t i = 0; i < D; i++) {
        sc sx; f32_to_scalar(sx, h_values[i], n_bits, frac);
        float xq = scalar_to_f32(sx, n_bits, frac), term = xq * xq * shift;
        val_float = (i == 0) ? term : val_float + term;
    }
    const float vf = scalar_to_f32(val, n_bits, frac);
    if (fabsf(vf - val_float) > 1.1920929e-7f) return 3;
    if (vf > l2_clip_max_f(range, n_bits, frac)) return 4;
    if (!(range == 8 || range == 16 || range == 32 || range == 64)) return -1;
    uint64_t v = (((uint64_t)val.v[1] << 32) | val.v[0]) & fix_max(n_bits);      // :69-73 read_from_bytes
    // V = v B + bsum H
    dev_buf d_v(8, s), d_bl(sizeof(sc_st), s), d_vs(sizeof(sc_st), s), d_V(32, s);
    sc vs; sc_from_u64(vs, v); sc_st t; sc_to_st(t, vs); rt_h2d(d_vs.p, &t, sizeof(t), s);
    sc_to_st(t, bsum); rt_h2d(d_bl.p, &t, sizeof(t), s); rt_h2d(d_v.p, &v, 8, s);
    { finalize_args f = {}; f.sBa = d_vs.as<sc_st>(); f.sHa = d_bl.as<sc_st>(); f.tabB = e.tabB; f.tabH = e.tabH; f.out32 = d_V.as<uint8_t>(); f.count = 1;
      run_finalize(s, f); }
    gens_entry &g = engine_gens(e, range, 1);                                     // BulletproofGens::new(64, 1) restricted to n = range (:162)
    std::vector<uint8_t> keys(32); derive_key(keys.data(), seed, DOM_L2_PROVE, 0);
    prove_chunks(e, s, "L2RangeProof", range, 1, 1, g, nullptr, d_v.as<uint64_t>(), d_bl.as<sc_st>(), d_V.as<uint8_t>(), keys, h_proof);
    rt_d2h(h_commit, d_V.p, 32, s); rt_sync(s);
    *proof_len = 32 * (9 + 2 * (size_t)ilog2_sz((size_t)range));
    return 0;
}
// verify_rangeproof_l2 (:185-228)
static int engine_l2_verify(rofl_engine &e, const uint8_t *h_proof, size_t plen, const uint8_t *h_commit, int range, const uint8_t seed[32]) {
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    if (plen % 32 || plen < 7 * 32) return -1;
    { size_t ne = plen / 32 - 7; if (ne < 2 || (ne - 2) % 2 || (ne - 2) / 2 >= 32) return -1; }
    dev_buf d_c(32, s), d_Vp3(sizeof(p3_st), s), d_V32(32, s), d_bad(sizeof(int), s);
    rt_h2d(d_c.p, h_commit, 32, s); rt_memset(d_bad.p, 0, sizeof(int), s);
    LAUNCH(k_decompress, dim3(1), dim3(128), s, d_Vp3.as<p3_st>(), d_V32.as<uint8_t>(), d_c.as<uint8_t>(), (size_t)1, (size_t)1, (const p3_st *)nullptr, d_bad.as<int>(), (size_t)1);
    uint8_t hV[32]; int bad = 0; rt_d2h(hV, d_V32.p, 32, s); rt_d2h(&bad, d_bad.p, sizeof(int), s); rt_sync(s);
    if (bad) return -4;
    if (!(range == 8 || range == 16 || range == 32 || range == 64)) {
        size_t lg = (plen / 32 - 9) / 2; const uint8_t *p = h_proof; sc t;
        for (size_t off : {(size_t)128, (size_t)160, (size_t)192, 224 + 64 * lg, 224 + 64 * lg + 32}) { sc_frombytes(t, p + off); if (!sc_is_canonical(t)) return -1; }
        return -2;
    }
    gens_entry &g = engine_gens(e, range, 1);
    std::vector<uint8_t> keys(32); derive_key(keys.data(), seed, DOM_L2_VERIFY, 0);
    std::vector<int> verdict;
    int rc = verify_chunks(e, s, "L2RangeProof", range, 1, 1, g, nullptr, d_Vp3.as<p3_st>(), hV, h_proof, plen, keys, verdict);
    if (rc < 0) return rc;
    return verdict.empty() ? 0 : verdict[0];
}

// =============================================================================================================================
// square_proof_vec::create_l2rangeproof_vec_existing / verify_l2rangeproof_vec (square_proof_vec/mod.rs:19-75,130-160); device pointers
// =============================================================================================================================
static int engine_square_prove(rofl_engine &e, const float *d_values, const uint8_t *d_value_com, const uint8_t *d_r1, const uint8_t *d_r2, size_t D,
                               int n_bits, int frac, const uint8_t seed[32], uint8_t *d_proofs, uint8_t *d_commits) {
    if (!fp_ok(n_bits, frac)) return -2;
    if (D == 0) return 0;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    dev_buf d_flags(sizeof(int), s); rt_memset(d_flags.p, 0, sizeof(int), s);
    square_args a = {}; a.values = d_values; a.value_com = d_value_com; a.r1 = d_r1; a.r2 = d_r2; a.D = D; a.n_bits = n_bits; a.frac = frac;
    uint8_t key[32]; derive_key(key, seed, DOM_SQUARE, 0); key_words(a.key, key);
    a.tabB = e.tabB; a.tabH = e.tabH; a.proofs = d_proofs; a.commits = d_commits; a.flags = d_flags.as<int>();
    void *tk = rt_prof_begin(PROF_SQUARE, s);
    LAUNCH(k_square_prove, dim3((unsigned)((D + 127) / 128)), dim3(128), s, a);
    rt_prof_end(PROF_SQUARE, tk, s);
    int flags = 0; rt_d2h(&flags, d_flags.p, sizeof(int), s); rt_sync(s);
    if (flags & 4) return -4;
    if (flags & 1) return -98;
    return 0;
}
static int engine_square_verify(rofl_engine &e, const uint8_t *d_proofs, const uint8_t *d_commits, size_t D) {
    if (D == 0) return 1;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    dev_buf d_res(2 * sizeof(int), s); int init[2] = {1, 0}; rt_h2d(d_res.p, init, sizeof(init), s);
    void *tk = rt_prof_begin(PROF_SQUARE, s);
    LAUNCH(k_square_verify, dim3((unsigned)((D + 127) / 128)), dim3(128), s, d_proofs, d_commits, D, e.tabB, e.tabH, d_res.as<int>());
    rt_prof_end(PROF_SQUARE, tk, s);
    int res[2]; rt_d2h(res, d_res.p, sizeof(res), s); rt_sync(s);
    if (res[1]) return -1;
    return res[0] ? 1 : 0;
}

// =============================================================================================================================
// compressed_rand_proof::CompressedRandProof::{helper_prove, helper_prove_existing, helper_verify} (compressed_rand_proof/mod.rs:134-158)
//   d_values / d_blind / d_value_com (nullable: L_i = commit(m_i, r_i)) are device pointers; h_proof (128 B) and h_pairs (D x 64) host.
//   returns 0 ok, -4 undecodable existing commitment, -6 more than 900 000 pairs (the reference's label table ends there and it panics), -98 NaN
// =============================================================================================================================
#define CRP_MAX_D 900000
static inline void crp_challenge(sc &c, const uint8_t *h_pairs, size_t D, const uint8_t cprime[64]) {
    transcript t; transcript_init(t, "CompressedRandProof");
    const uint8_t ds[19] = {'r', 'a', 'n', 'd', 'o', 'm', 'n', 'e', 's', 's', ' ', 'p', 'r', 'o', 'o', 'f', ' ', 'v', '1'};
    transcript_append(t, "dom-sep", ds, 19);                                                         // rand_proof/transcript.rs:20-22
    for (size_t i = 0; i < D; i++) {                                                                 // dealer.rs:27-29; labels: generate_unique_u8_triplets.py:9-13
        const uint8_t lab[3] = {(uint8_t)(3 * i), (uint8_t)(3 * i + 1), (uint8_t)(3 * i + 2)};
        transcript_append_l(t, lab, 3, h_pairs + 64 * i, 64);
    }
    transcript_append(t, "C_prime_eg", cprime, 64);                                                  // dealer.rs:53-54
    ts_challenge_scalar(t, "c", c);
}
static inline pow_tab crp_pow_table(rofl_engine &e, cudaStream_t s, const sc &c, size_t D, dev_buf &store) {
    const int bits = std::max(2, ilog2_sz(D + 2));
    pow_tab ct = {nullptr, bits / 2, bits - bits / 2};
    std::vector<sc_st> h_pow2(32); sc_pow2_table(h_pow2.data(), c);
    sc_st *base = store.as<sc_st>();                               // [32 pow2 | table]
    rt_h2d(base, h_pow2.data(), sizeof(sc_st) * 32, s);
    ct.tab = base + 32;
    LAUNCH(k_pow_tables, dim3((pow_tab_size(ct) + 255) / 256, 1), dim3(256), s, base + 32, base, ct.L, ct.H);
    return ct;
}
static int engine_crp_prove(rofl_engine &e, const float *d_values, const uint8_t *d_value_com, const uint8_t *d_blind, size_t D, int n_bits, int frac,
                            const uint8_t seed[32], uint8_t *h_proof, uint8_t *h_pairs) {
    if (!fp_ok(n_bits, frac)) return -2;
    if (D > CRP_MAX_D) return -6;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    const size_t Dn = D ? D : 1;
    dev_buf d_L(32 * Dn, s), d_R(32 * Dn, s), d_pairs(64 * Dn, s), d_flags(sizeof(int), s), d_bad(sizeof(int), s);
    rt_memset(d_flags.p, 0, sizeof(int), s); rt_memset(d_bad.p, 0, sizeof(int), s);
    if (D) {
        commit_args ca = {}; ca.values = d_values; ca.blind = d_blind; ca.D = D; ca.Dp = D; ca.n_bits = n_bits; ca.frac = frac;
        ca.tabB = e.tabB; ca.tabH = e.tabH; ca.C = d_value_com ? nullptr : d_L.as<uint8_t>(); ca.R = d_R.as<uint8_t>(); ca.flags = d_flags.as<int>();   // R_i = r_i B (el_gamal.rs:57-69)
        LAUNCH(k_commit, dim3((unsigned)((D + 127) / 128)), dim3(128), s, ca);
        if (d_value_com) {          // prove_existing: the caller's commitments must at least decode (the reference holds RistrettoPoints)
            dev_buf d_tmp(sizeof(p3_st) * D, s);
            LAUNCH(k_decompress, dim3((unsigned)((D + 127) / 128)), dim3(128), s, d_tmp.as<p3_st>(), (uint8_t *)nullptr, d_value_com, D, D, (const p3_st *)nullptr, d_bad.as<int>(), D);
        }
        LAUNCH(k_pairs_join, dim3((unsigned)((D + 255) / 256)), dim3(256), s, d_pairs.as<uint8_t>(), d_value_com ? d_value_com : d_L.as<uint8_t>(), d_R.as<uint8_t>(), D);
        rt_d2h(h_pairs, d_pairs.p, 64 * D, s);
    }
    // C' = (m' B + r' H, r' B), nonces in the draw order of Party::new (party.rs:26-29)
    uint8_t key[32]; derive_key(key, seed, DOM_CRP, 0); uint32_t kw[8]; key_words(kw, key);
    sc mp, rp, zero; nonce_scalar(mp, kw, 0); nonce_scalar(rp, kw, 1); sc_0(zero);
    sc_st h_sB[2], h_sH[2]; sc_to_st(h_sB[0], mp); sc_to_st(h_sH[0], rp); sc_to_st(h_sB[1], rp); sc_to_st(h_sH[1], zero);
    dev_buf d_s(sizeof(sc_st) * 4, s), d_cp(64, s);
    rt_h2d(d_s.p, h_sB, sizeof(h_sB), s); rt_h2d(d_s.as<sc_st>() + 2, h_sH, sizeof(h_sH), s);
    { finalize_args f = {}; f.sBa = d_s.as<sc_st>(); f.sHa = d_s.as<sc_st>() + 2; f.tabB = e.tabB; f.tabH = e.tabH; f.out32 = d_cp.as<uint8_t>(); f.count = 2; run_finalize(s, f); }
    int flags = 0, bad = 0;
    rt_d2h(h_proof, d_cp.p, 64, s); rt_d2h(&flags, d_flags.p, sizeof(int), s); rt_d2h(&bad, d_bad.p, sizeof(int), s);
    rt_sync(s);
    if (flags & 1) return -98;
    if (bad) return -4;
    sc c; crp_challenge(c, h_pairs, D, h_proof);
    sc zm = mp, zr = rp;
    if (D) {
        dev_buf d_ct(sizeof(sc_st) * (32 + ((size_t)2 << ((ilog2_sz(D + 2) + 1) / 2 + 1))), s);
        pow_tab ct = crp_pow_table(e, s, c, D, d_ct);
        const int nb = (int)std::min<size_t>(256, (D + 255) / 256);
        dev_buf d_part(sizeof(sc_st) * 2 * nb, s), d_sum(sizeof(sc_st) * 2, s);
        LAUNCH_COOP(k_crp_sums, dim3(nb), dim3(256), s, d_part.as<sc_st>(), d_values, d_blind, D, n_bits, frac, ct, d_flags.as<int>());
        LAUNCH_COOP(k_sc_sum, dim3(1), dim3(256), s, d_sum.as<sc_st>(), d_part.as<sc_st>(), nb, 2);
        sc_st hs[2]; rt_d2h(hs, d_sum.p, sizeof(hs), s); rt_sync(s);
        sc a, b; st_to_sc(a, hs[0]); st_to_sc(b, hs[1]); sc_add(zm, zm, a); sc_add(zr, zr, b);     // party.rs:93-97
    }
    sc_tobytes(h_proof + 64, zm); sc_tobytes(h_proof + 96, zr);
    return 0;
}
// returns 1 valid, 0 invalid, -1 FormatError (from_bytes: mod.rs:118-134, el_gamal.rs:113-123), -6 too many pairs
static int engine_crp_verify(rofl_engine &e, const uint8_t *h_proof, const uint8_t *h_pairs, size_t D) {
    if (D > CRP_MAX_D) return -6;
    sc zm, zr; sc_frombytes(zm, h_proof + 64); sc_frombytes(zr, h_proof + 96);
    if (!sc_is_canonical(zm) || !sc_is_canonical(zr)) return -1;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    const size_t Dn = D ? D : 1;
    dev_buf d_pairs(64 * Dn, s), d_LR32(64 * Dn, s), d_LR(sizeof(p3_st) * 2 * Dn, s), d_cp32(64, s), d_cp(sizeof(p3_st) * 2, s), d_bad(sizeof(int) * 2, s);
    rt_memset(d_bad.p, 0, sizeof(int) * 2, s);
    rt_h2d(d_cp32.p, h_proof, 64, s);
    LAUNCH(k_decompress, dim3(1), dim3(128), s, d_cp.as<p3_st>(), (uint8_t *)nullptr, d_cp32.as<uint8_t>(), (size_t)2, (size_t)2, (const p3_st *)nullptr, d_bad.as<int>(), (size_t)2);
    if (D) {
        rt_h2d(d_pairs.p, h_pairs, 64 * D, s);
        LAUNCH(k_pairs_split, dim3((unsigned)((D + 255) / 256)), dim3(256), s, d_LR32.as<uint8_t>(), d_LR32.as<uint8_t>() + 32 * D, d_pairs.as<uint8_t>(), D);
        LAUNCH(k_decompress, dim3((unsigned)((2 * D + 127) / 128)), dim3(128), s, d_LR.as<p3_st>(), (uint8_t *)nullptr, d_LR32.as<uint8_t>(), 2 * D, 2 * D, (const p3_st *)nullptr, d_bad.as<int>() + 1, 2 * D);
    }
    sc c; crp_challenge(c, h_pairs, D, h_proof);                   // overlaps the decompression
    // sum_i c^(i+1) L_i and sum_i c^(i+1) R_i: two MSMs that share their scalars
    sc nzm, nzr, zero; sc_neg(nzm, zm); sc_neg(nzr, zr); sc_0(zero);
    sc_st h_s[4]; sc_to_st(h_s[0], nzm); sc_to_st(h_s[1], nzr); sc_to_st(h_s[2], nzr); sc_to_st(h_s[3], zero);      // sB = [-z_m, -z_r], sH = [-z_r, 0]
    dev_buf d_s(sizeof(sc_st) * 4, s), d_id(sizeof(int) * 2, s);
    rt_h2d(d_s.p, h_s, sizeof(h_s), s);
    finalize_args f = {}; f.partial = d_cp.as<p3_st>(); f.npartial = 1; f.sBa = d_s.as<sc_st>(); f.sHa = d_s.as<sc_st>() + 2; f.tabB = e.tabB; f.tabH = e.tabH;
    f.is_id = d_id.as<int>(); f.count = 2;
    if (D) {
        dev_buf d_ct(sizeof(sc_st) * (32 + ((size_t)2 << ((ilog2_sz(D + 2) + 1) / 2 + 1))), s), d_pw(sizeof(sc_st) * D, s);
        pow_tab ct = crp_pow_table(e, s, c, D, d_ct);
        LAUNCH(k_crp_pows, dim3((unsigned)((D + 255) / 256)), dim3(256), s, d_pw.as<sc_st>(), D, ct);
        const msm_plan pl = msm_plan_for(D, 2);
        dev_buf d_win(sizeof(p3_st) * pl.out_count(2), s);
        msm_args a = {}; a.v[0].scalars = d_pw.as<sc_st>(); a.v[1].scalars = d_pw.as<sc_st>(); a.split = 1; a.T = (uint32_t)D; a.scalar_stride = 0; a.nseg = 1; a.out = d_win.as<p3_st>();
        a.v[0].seg[0] = mk_seg(d_LR.as<p3_st>(), (uint32_t)D, 0, 1); a.v[1].seg[0] = mk_seg(d_LR.as<p3_st>() + D, (uint32_t)D, 0, 1);
        run_msm(e, s, a, pl, 2);
        fin_windows(f, d_win.as<p3_st>(), pl);
        run_finalize(s, f);
        int id[2], bad[2]; rt_d2h(id, d_id.p, sizeof(id), s); rt_d2h(bad, d_bad.p, sizeof(bad), s); rt_sync(s);
        if (bad[0] || bad[1]) return -1;
        return (id[0] && id[1]) ? 1 : 0;
    }
    run_finalize(s, f);
    int id[2], bad[2]; rt_d2h(id, d_id.p, sizeof(id), s); rt_d2h(bad, d_bad.p, sizeof(bad), s); rt_sync(s);
    if (bad[0]) return -1;
    return (id[0] && id[1]) ? 1 : 0;
}

// rand_proof_vec::{create_randproof_vec(_existing), verify_randproof_vec} (kind 1) and square_rand_proof_vec::{create_l2rangeproof_vec(_existing),
// verify_l2rangeproof_vec} (kind 2); device pointers; d_value_com nullable (commit inside).  prove: 0, -4 bad point, -98 NaN; verify: 1 / 0 / -1 FormatError
static int engine_sigma_prove(rofl_engine &e, int kind, const float *d_values, const uint8_t *d_value_com, const uint8_t *d_r1, const uint8_t *d_r2, size_t D,
                              int n_bits, int frac, const uint8_t seed[32], uint8_t *d_proofs, uint8_t *d_commits) {
    if (!fp_ok(n_bits, frac) || (kind != 1 && kind != 2)) return -2;
    if (D == 0) return 0;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    dev_buf d_flags(sizeof(int), s); rt_memset(d_flags.p, 0, sizeof(int), s);
    sigma_args a = {}; a.kind = kind; a.values = d_values; a.value_com = d_value_com; a.r1 = d_r1; a.r2 = d_r2; a.D = D; a.n_bits = n_bits; a.frac = frac;
    uint8_t key[32]; derive_key(key, seed, kind == 1 ? DOM_RANDPROOF : DOM_SQUARE_RAND, 0); key_words(a.key, key);
    a.tabB = e.tabB; a.tabH = e.tabH; a.proofs = d_proofs; a.commits = d_commits; a.flags = d_flags.as<int>();
    void *tk = rt_prof_begin(PROF_SQUARE, s);
    LAUNCH(k_sigma_prove, dim3((unsigned)((D + 127) / 128)), dim3(128), s, a);
    rt_prof_end(PROF_SQUARE, tk, s);
    int flags = 0; rt_d2h(&flags, d_flags.p, sizeof(int), s); rt_sync(s);
    if (flags & 4) return -4;
    if (flags & 1) return -98;
    return 0;
}
static int engine_sigma_verify(rofl_engine &e, int kind, const uint8_t *d_proofs, const uint8_t *d_commits, size_t D) {
    if (kind != 1 && kind != 2) return -2;
    if (D == 0) return 1;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    dev_buf d_res(2 * sizeof(int), s); int init[2] = {1, 0}; rt_h2d(d_res.p, init, sizeof(init), s);
    LAUNCH(k_sigma_verify, dim3((unsigned)((D + 127) / 128)), dim3(128), s, kind, d_proofs, d_commits, D, e.tabB, e.tabH, d_res.as<int>());
    int res[2]; rt_d2h(res, d_res.p, sizeof(res), s); rt_sync(s);
    if (res[1]) return -1;
    return res[0] ? 1 : 0;
}

// sum of D compressed points -> compressed (device in, host out); -4 if one does not decode
static int engine_points_sum(rofl_engine &e, const uint8_t *d_pts, size_t D, uint8_t *h_out32) {
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    const int nb = (int)std::max<size_t>(1, std::min<size_t>(296, (D + 127) / 128));
    dev_buf d_part(sizeof(p3_st) * nb, s), d_bad(sizeof(int), s), d_out(32, s);
    rt_memset(d_bad.p, 0, sizeof(int), s);
    LAUNCH_COOP(k_points_sum, dim3(nb), dim3(128), s, d_part.as<p3_st>(), d_pts, D, d_bad.as<int>());
    finalize_args f = {}; f.partial = d_part.as<p3_st>(); f.npartial = nb; f.tabB = e.tabB; f.tabH = e.tabH; f.out32 = d_out.as<uint8_t>(); f.count = 1;
    run_finalize(s, f);
    int bad = 0; rt_d2h(h_out32, d_out.p, 32, s); rt_d2h(&bad, d_bad.p, sizeof(int), s); rt_sync(s);
    return bad ? -4 : 0;
}

// =============================================================================================================================
// commitments (pedersen_ops.rs:9-25; el_gamal.rs:57-69), aggregation (params.rs:81-124), discrete log (bsgs32.rs, pedersen_ops.rs:47-53)
// =============================================================================================================================
static int engine_commit(rofl_engine &e, const float *d_values, const uint8_t *d_blind, size_t D, int n_bits, int frac, uint8_t *d_L, uint8_t *d_R) {
    if (!fp_ok(n_bits, frac)) return -2;
    if (D == 0) return 0;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    dev_buf d_flags(sizeof(int), s); rt_memset(d_flags.p, 0, sizeof(int), s);
    commit_args ca = {}; ca.values = d_values; ca.blind = d_blind; ca.D = D; ca.Dp = D; ca.n_bits = n_bits; ca.frac = frac;
    ca.tabB = e.tabB; ca.tabH = e.tabH; ca.C = d_L; ca.R = d_blind ? d_R : nullptr; ca.flags = d_flags.as<int>();
    void *tk = rt_prof_begin(PROF_COMMIT, s);
    LAUNCH(k_commit, dim3((unsigned)((D + 127) / 128)), dim3(128), s, ca);
    rt_prof_end(PROF_COMMIT, tk, s);
    int flags = 0; rt_d2h(&flags, d_flags.p, sizeof(int), s); rt_sync(s);
    return (flags & 1) ? -98 : 0;
}
static int engine_aggregate(rofl_engine &e, const uint8_t *d_pts, size_t n_clients, size_t D, int init_base, uint8_t *d_out) {
    if (D == 0) return 0;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    dev_buf d_bad(sizeof(int), s); rt_memset(d_bad.p, 0, sizeof(int), s);
    LAUNCH(k_aggregate, dim3((unsigned)((D + 127) / 128)), dim3(128), s, d_out, d_pts, n_clients, D, init_base, d_bad.as<int>());
    int bad = 0; rt_d2h(&bad, d_bad.p, sizeof(int), s); rt_sync(s);
    return bad ? -4 : 0;
}
static bsgs_entry &engine_bsgs(rofl_engine &e, uint64_t table_size, int bsgs_bits) {
    auto key = std::make_pair(table_size, bsgs_bits);
    auto it = e.bsgs.find(key);
    if (it != e.bsgs.end()) return it->second;
    cudaStream_t s = e.stream;
    bsgs_entry b; b.cap = 1; while (b.cap < 4 * (table_size + 1)) b.cap <<= 1;
    b.keys = (unsigned long long *)rt_malloc(8 * (size_t)b.cap, s); b.vals = (uint32_t *)rt_malloc(4 * (size_t)b.cap, s);
    rt_memset(b.keys, 0xff, 8 * (size_t)b.cap, s); rt_memset(b.vals, 0, 4 * (size_t)b.cap, s);
    dev_buf d_cnt(sizeof(int), s); rt_memset(d_cnt.p, 0, sizeof(int), s);
    LAUNCH(k_bsgs_build, dim3((unsigned)((table_size + 1 + 127) / 128)), dim3(128), s, b.keys, b.vals, b.cap, (uint32_t)table_size, e.tabB, d_cnt.as<int>());
    int cnt = 0; rt_d2h(&cnt, d_cnt.p, sizeof(int), s); rt_sync(s);
    b.size = (uint64_t)cnt - 1;                                                   // get_size (bsgs32.rs:44-46)
    return e.bsgs[key] = b;
}
// returns 0, -4 undecodable point, -5 where the reference panics (no discrete log within range for either sign, bsgs32.rs:69-70)
static int engine_dlog(rofl_engine &e, const uint8_t *d_pts, size_t D, uint64_t table_size, int bsgs_bits, int n_bits, int frac, uint8_t *d_out_sc, float *d_out_f32) {
    if (table_size == 0 || table_size > (1ull << 28) || bsgs_bits < 1 || bsgs_bits > 32) return -2;
    if (D == 0) return 0;
    std::lock_guard<std::mutex> lk(e.mu);
    cudaStream_t s = e.stream;
    bsgs_entry &b = engine_bsgs(e, table_size, bsgs_bits);
    uint64_t max_it = b.size ? (1ULL << bsgs_bits) / b.size : 0;                  // bsgs32.rs:60-62
    dev_buf d_flags(sizeof(int), s); rt_memset(d_flags.p, 0, sizeof(int), s);
    LAUNCH(k_bsgs_solve, dim3((unsigned)((D + 127) / 128)), dim3(128), s, d_out_sc, d_out_f32, d_pts, D, b.keys, b.vals, b.cap, (uint32_t)table_size, b.size, max_it,
           bsgs_bits, n_bits, frac, e.tabB, d_flags.as<int>());
    int flags = 0; rt_d2h(&flags, d_flags.p, sizeof(int), s); rt_sync(s);
    if (flags & 4) return -4;
    if (flags & 8) return -5;
    return 0;
}
