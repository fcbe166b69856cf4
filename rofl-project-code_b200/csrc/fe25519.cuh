// GF(2^255-19) for sm_100a: 8 saturated 32-bit limbs.  Every element is ANY 256-bit integer (a "loose" representative
// mod p; 2^256 = 38 mod p), so there are no limb-scale rules: every function accepts and returns loose values.
//
// Device code is column-wise (Comba) multiplication in inline PTX: each of the 15 product columns is an independent
// 96-bit accumulator (lo, hi, ex) fed by `mad.lo.cc / madc.hi.cc / addc`, which ptxas fuses into ONE
// `IMAD.WIDE.U32 Rd, Pc, Ra, Rb, Rd` (64-bit accumulate with carry-out predicate) plus half an `IADD3.X` per product
// (tools/microbench2.cu measures that pair at the IMAD.WIDE issue rate).  A multiplication is 64 + 8 IMAD.WIDE.U32 and a
// squaring 36 + 8, against 100 / 55 for the radix-2^25.5 representation this replaces; the IMAD.WIDE pipe
// (32 lanes/clk/SM) is the roofline of the whole prove / verify path.
//
// This is the arithmetic the reference reaches through curve25519-dalek-ng `FieldElement` (rofl_crypto/Cargo.toml:13;
// third-party).  Written from the field definition; all functions are __host__ __device__ with a portable C path for the
// host so tests/hostsim can run the same point / protocol code on the CPU against the oracle.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDNI static __host__ __device__ __noinline__
#else
#define HD inline
#define HDNI static inline
#endif
#define FE_ASSERT(c) ((void)0)
// fe_mul / fe_sq are inlined into the point formulas.  -DFE_INLINE_MUL=0 turns them into real device functions (smaller code,
// friendlier to the instruction cache) but that build is NOT bit-exact on sm_100a with CUDA 12.9: tools/ge_selftest.cu shows a
// doubling + addition + compress sequence going wrong only in that configuration (no memcheck / racecheck findings), so it
// stays off until the cause is understood.  The GPU parity tests would catch the same failure in any other build.
#ifndef FE_MUL_ROWS
#define FE_MUL_ROWS 1
#endif
#ifndef FE_INLINE_MUL
#define FE_INLINE_MUL 1
#endif
#if defined(__CUDACC__) && !FE_INLINE_MUL
#define HDMUL static __host__ __device__ __noinline__
#elif defined(__CUDACC__)
#define HDMUL __host__ __device__ __forceinline__
#else
#define HDMUL inline
#endif

#if !defined(__CUDACC__)
// host-only builds (tests/hostsim): the few CUDA vector types the storage helpers use
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define __align__(n) alignas(n)
#endif

#define FE_LIMBS 8
struct fe { uint32_t v[8]; };

HD void fe_0(fe &h) { for (int i = 0; i < 8; i++) h.v[i] = 0; }
HD void fe_1(fe &h) { fe_0(h); h.v[0] = 1; }
HD void fe_carry(fe &h, const fe &f) { h = f; }          // kept for source compatibility: loose values need no carry pass

// h = f + g: 256-bit add; a carry out of bit 256 is worth 38
HD void fe_add(fe &h, const fe &f, const fe &g) {
#if defined(__CUDA_ARCH__)
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7, c;
    asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\taddc.u32 %8, 0, 0;"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3), "=&r"(r4), "=&r"(r5), "=&r"(r6), "=&r"(r7), "=&r"(c)
        : "r"(f.v[0]), "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[6]), "r"(f.v[7]),
          "r"(g.v[0]), "r"(g.v[1]), "r"(g.v[2]), "r"(g.v[3]), "r"(g.v[4]), "r"(g.v[5]), "r"(g.v[6]), "r"(g.v[7]));
    c *= 38;
    asm("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.cc.u32 %6, %6, 0;\n\taddc.cc.u32 %7, %7, 0;\n\taddc.u32 %8, 0, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(c));
    r0 += 38 * c;                                        // second wrap only when the sum was >= 2^256 - 38: r0 < 38, no carry
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
    uint64_t c = 0; uint32_t r[8];
    for (int i = 0; i < 8; i++) { c += (uint64_t)f.v[i] + g.v[i]; r[i] = (uint32_t)c; c >>= 32; }
    c *= 38;
    for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
    r[0] += 38 * (uint32_t)c;
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
#endif
}
// h = f - g: 256-bit subtract; a borrow out of bit 256 is worth -38
HD void fe_sub(fe &h, const fe &f, const fe &g) {
#if defined(__CUDA_ARCH__)
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7, b;
    asm("sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\tsubc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\tsubc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\tsubc.u32 %8, 0, 0;"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3), "=&r"(r4), "=&r"(r5), "=&r"(r6), "=&r"(r7), "=&r"(b)
        : "r"(f.v[0]), "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[6]), "r"(f.v[7]),
          "r"(g.v[0]), "r"(g.v[1]), "r"(g.v[2]), "r"(g.v[3]), "r"(g.v[4]), "r"(g.v[5]), "r"(g.v[6]), "r"(g.v[7]));
    b = (b & 1) * 38;                                    // b = 0xffffffff on borrow
    asm("sub.cc.u32 %0, %0, %8;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.cc.u32 %2, %2, 0;\n\tsubc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\tsubc.cc.u32 %5, %5, 0;\n\tsubc.cc.u32 %6, %6, 0;\n\tsubc.cc.u32 %7, %7, 0;\n\tsubc.u32 %8, 0, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(b));
    r0 -= 38 * (b & 1);                                  // second wrap only when the difference was < 38 - 2^256 ...: r0 >= 2^32 - 38, no borrow
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
    int64_t c = 0; uint32_t r[8];
    for (int i = 0; i < 8; i++) { c += (int64_t)f.v[i] - (int64_t)g.v[i]; r[i] = (uint32_t)c; c >>= 32; }
    c *= 38;                                             // c = 0 or -1
    for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
    r[0] -= 38 * (uint32_t)(-c);
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
#endif
}
HD void fe_sub4(fe &h, const fe &f, const fe &g) { fe_sub(h, f, g); }
HD void fe_neg(fe &h, const fe &f) { fe z; fe_0(z); fe_sub(h, z, f); }

// ---- reduction of a 512-bit product t[0..15] to a loose 256-bit value: lo + 38 hi, then fold the 6-bit overflow twice ----
#if defined(__CUDA_ARCH__)
// d = a*b + c (64-bit, cannot overflow for b = 38): one IMAD.WIDE.U32
__device__ __forceinline__ uint64_t fe_mad_wide(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"((uint64_t)c)); return d;
}
__device__ __forceinline__ void fe_reduce512(fe &h, const uint32_t t[16]) {
    uint64_t d0 = fe_mad_wide(t[8], 38, t[0]), d1 = fe_mad_wide(t[9], 38, t[1]), d2 = fe_mad_wide(t[10], 38, t[2]), d3 = fe_mad_wide(t[11], 38, t[3]);
    uint64_t d4 = fe_mad_wide(t[12], 38, t[4]), d5 = fe_mad_wide(t[13], 38, t[5]), d6 = fe_mad_wide(t[14], 38, t[6]), d7 = fe_mad_wide(t[15], 38, t[7]);
    uint32_t r0 = (uint32_t)d0, r1, r2, r3, r4, r5, r6, r7, top;
    asm("add.cc.u32 %0, %8, %9;\n\taddc.cc.u32 %1, %10, %11;\n\taddc.cc.u32 %2, %12, %13;\n\taddc.cc.u32 %3, %14, %15;\n\t"
        "addc.cc.u32 %4, %16, %17;\n\taddc.cc.u32 %5, %18, %19;\n\taddc.cc.u32 %6, %20, %21;\n\taddc.u32 %7, %22, 0;"
        : "=&r"(r1), "=&r"(r2), "=&r"(r3), "=&r"(r4), "=&r"(r5), "=&r"(r6), "=&r"(r7), "=&r"(top)
        : "r"((uint32_t)d1), "r"((uint32_t)(d0 >> 32)), "r"((uint32_t)d2), "r"((uint32_t)(d1 >> 32)), "r"((uint32_t)d3), "r"((uint32_t)(d2 >> 32)),
          "r"((uint32_t)d4), "r"((uint32_t)(d3 >> 32)), "r"((uint32_t)d5), "r"((uint32_t)(d4 >> 32)), "r"((uint32_t)d6), "r"((uint32_t)(d5 >> 32)),
          "r"((uint32_t)d7), "r"((uint32_t)(d6 >> 32)), "r"((uint32_t)(d7 >> 32)));
    uint32_t c = top * 38;                               // top <= 38
    asm("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.cc.u32 %6, %6, 0;\n\taddc.cc.u32 %7, %7, 0;\n\taddc.u32 %8, 0, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(c));
    r0 += 38 * c;
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
}
// column accumulator: (lo, hi, ex) += a*b
#define FE_MAC(lo, hi, ex, a, b) asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(ex) : "r"(a), "r"(b))
#define FE_MUL0(lo, hi, a, b) do { uint64_t p_; asm("mul.wide.u32 %0, %1, %2;" : "=l"(p_) : "r"(a), "r"(b)); lo = (uint32_t)p_; hi = (uint32_t)(p_ >> 32); } while (0)
// fold 15 column accumulators into 16 words: word k = lo_k + hi_(k-1) + ex_(k-2) + carries
__device__ __forceinline__ void fe_columns_to_words(uint32_t t[16], const uint32_t lo[15], const uint32_t hi[15], const uint32_t ex[15]) {
    t[0] = lo[0];
    asm("add.cc.u32 %0, %15, %30;\n\taddc.cc.u32 %1, %16, %31;\n\taddc.cc.u32 %2, %17, %32;\n\taddc.cc.u32 %3, %18, %33;\n\taddc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\taddc.cc.u32 %6, %21, %36;\n\taddc.cc.u32 %7, %22, %37;\n\taddc.cc.u32 %8, %23, %38;\n\taddc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\taddc.cc.u32 %11, %26, %41;\n\taddc.cc.u32 %12, %27, %42;\n\taddc.cc.u32 %13, %28, %43;\n\taddc.u32 %14, %29, 0;"
        : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8]), "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11]), "=&r"(t[12]), "=&r"(t[13]), "=&r"(t[14]), "=&r"(t[15])
        : "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(lo[8]), "r"(lo[9]), "r"(lo[10]), "r"(lo[11]), "r"(lo[12]), "r"(lo[13]), "r"(lo[14]), "r"(hi[14]),
          "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]), "r"(hi[8]), "r"(hi[9]), "r"(hi[10]), "r"(hi[11]), "r"(hi[12]), "r"(hi[13]));
    asm("add.cc.u32 %0, %0, %14;\n\taddc.cc.u32 %1, %1, %15;\n\taddc.cc.u32 %2, %2, %16;\n\taddc.cc.u32 %3, %3, %17;\n\taddc.cc.u32 %4, %4, %18;\n\t"
        "addc.cc.u32 %5, %5, %19;\n\taddc.cc.u32 %6, %6, %20;\n\taddc.cc.u32 %7, %7, %21;\n\taddc.cc.u32 %8, %8, %22;\n\taddc.cc.u32 %9, %9, %23;\n\t"
        "addc.cc.u32 %10, %10, %24;\n\taddc.cc.u32 %11, %11, %25;\n\taddc.cc.u32 %12, %12, %26;\n\taddc.u32 %13, %13, %27;"
        : "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
        : "r"(ex[0]), "r"(ex[1]), "r"(ex[2]), "r"(ex[3]), "r"(ex[4]), "r"(ex[5]), "r"(ex[6]), "r"(ex[7]), "r"(ex[8]), "r"(ex[9]), "r"(ex[10]), "r"(ex[11]), "r"(ex[12]), "r"(ex[13]));
}
// h = f*g, row-wise: products of equal parity of (i + j) accumulate into 64-bit aligned register pairs along a carry chain
// (IMAD.WIDE.U32 with carry-in and carry-out), so no per-product carry word is needed; E holds the pairs at even word
// positions, O those at odd positions, T = E + O.  Chain ends: the running sum after row i is < 2^(32(i+9)), so a chain that
// ends at word i+8 cannot carry out and one that ends at word i+7 carries into a still-untouched word.
__device__ __forceinline__ void fe_mul_rows(uint32_t t[16], const fe &f, const fe &g) {
    uint32_t e[16], o[17];
    for (int k = 0; k < 16; k++) e[k] = 0;
    for (int k = 0; k < 17; k++) o[k] = 0;
    FE_MUL0(e[0], e[1], f.v[0], g.v[0]);
    FE_MUL0(e[2], e[3], f.v[0], g.v[2]);
    FE_MUL0(e[4], e[5], f.v[0], g.v[4]);
    FE_MUL0(e[6], e[7], f.v[0], g.v[6]);
    FE_MUL0(o[1], o[2], f.v[0], g.v[1]);
    FE_MUL0(o[3], o[4], f.v[0], g.v[3]);
    FE_MUL0(o[5], o[6], f.v[0], g.v[5]);
    FE_MUL0(o[7], o[8], f.v[0], g.v[7]);
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\tmadc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\tmadc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
        : "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9])
        : "r"(f.v[1]), "r"(g.v[1]), "r"(g.v[3]), "r"(g.v[5]), "r"(g.v[7]));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\tmadc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\tmadc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9])
        : "r"(f.v[1]), "r"(g.v[0]), "r"(g.v[2]), "r"(g.v[4]), "r"(g.v[6]));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\tmadc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\tmadc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "+r"(e[10])
        : "r"(f.v[2]), "r"(g.v[0]), "r"(g.v[2]), "r"(g.v[4]), "r"(g.v[6]));
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\tmadc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\tmadc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
        : "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10])
        : "r"(f.v[2]), "r"(g.v[1]), "r"(g.v[3]), "r"(g.v[5]), "r"(g.v[7]));
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\tmadc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\tmadc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
        : "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "+r"(e[10]), "+r"(e[11])
        : "r"(f.v[3]), "r"(g.v[1]), "r"(g.v[3]), "r"(g.v[5]), "r"(g.v[7]));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\tmadc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\tmadc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11])
        : "r"(f.v[3]), "r"(g.v[0]), "r"(g.v[2]), "r"(g.v[4]), "r"(g.v[6]));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\tmadc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\tmadc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "+r"(e[10]), "+r"(e[11]), "+r"(e[12])
        : "r"(f.v[4]), "r"(g.v[0]), "r"(g.v[2]), "r"(g.v[4]), "r"(g.v[6]));
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\tmadc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\tmadc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
        : "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11]), "+r"(o[12])
        : "r"(f.v[4]), "r"(g.v[1]), "r"(g.v[3]), "r"(g.v[5]), "r"(g.v[7]));
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\tmadc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\tmadc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
        : "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "+r"(e[10]), "+r"(e[11]), "+r"(e[12]), "+r"(e[13])
        : "r"(f.v[5]), "r"(g.v[1]), "r"(g.v[3]), "r"(g.v[5]), "r"(g.v[7]));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\tmadc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\tmadc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11]), "+r"(o[12]), "+r"(o[13])
        : "r"(f.v[5]), "r"(g.v[0]), "r"(g.v[2]), "r"(g.v[4]), "r"(g.v[6]));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\tmadc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\tmadc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "+r"(e[10]), "+r"(e[11]), "+r"(e[12]), "+r"(e[13]), "+r"(e[14])
        : "r"(f.v[6]), "r"(g.v[0]), "r"(g.v[2]), "r"(g.v[4]), "r"(g.v[6]));
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\tmadc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\tmadc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
        : "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11]), "+r"(o[12]), "+r"(o[13]), "+r"(o[14])
        : "r"(f.v[6]), "r"(g.v[1]), "r"(g.v[3]), "r"(g.v[5]), "r"(g.v[7]));
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\tmadc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\tmadc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
        : "+r"(e[8]), "+r"(e[9]), "+r"(e[10]), "+r"(e[11]), "+r"(e[12]), "+r"(e[13]), "+r"(e[14]), "+r"(e[15])
        : "r"(f.v[7]), "r"(g.v[1]), "r"(g.v[3]), "r"(g.v[5]), "r"(g.v[7]));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\tmadc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\tmadc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\tmadc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\taddc.u32 %8, 0, 0;"
        : "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11]), "+r"(o[12]), "+r"(o[13]), "+r"(o[14]), "+r"(o[15])
        : "r"(f.v[7]), "r"(g.v[0]), "r"(g.v[2]), "r"(g.v[4]), "r"(g.v[6]));
    t[0] = e[0];
    asm("add.cc.u32 %0, %15, %30;\n\taddc.cc.u32 %1, %16, %31;\n\taddc.cc.u32 %2, %17, %32;\n\taddc.cc.u32 %3, %18, %33;\n\taddc.cc.u32 %4, %19, %34;\n\taddc.cc.u32 %5, %20, %35;\n\taddc.cc.u32 %6, %21, %36;\n\taddc.cc.u32 %7, %22, %37;\n\taddc.cc.u32 %8, %23, %38;\n\taddc.cc.u32 %9, %24, %39;\n\taddc.cc.u32 %10, %25, %40;\n\taddc.cc.u32 %11, %26, %41;\n\taddc.cc.u32 %12, %27, %42;\n\taddc.cc.u32 %13, %28, %43;\n\taddc.u32 %14, %29, %44;"
        : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8]), "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11]), "=&r"(t[12]), "=&r"(t[13]), "=&r"(t[14]), "=&r"(t[15])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]));
}
// off-diagonal products f_i f_j (i < j) of a square, row-wise as in fe_mul_rows: t = sum_{i<j} f_i f_j 2^(32(i+j))
__device__ __forceinline__ void fe_sq_rows(uint32_t t[16], const fe &f) {
    uint32_t e[16], o[17];
    for (int k = 0; k < 16; k++) e[k] = 0;
    for (int k = 0; k < 17; k++) o[k] = 0;
    FE_MUL0(e[2], e[3], f.v[0], f.v[2]);
    FE_MUL0(e[4], e[5], f.v[0], f.v[4]);
    FE_MUL0(e[6], e[7], f.v[0], f.v[6]);
    FE_MUL0(o[1], o[2], f.v[0], f.v[1]);
    FE_MUL0(o[3], o[4], f.v[0], f.v[3]);
    FE_MUL0(o[5], o[6], f.v[0], f.v[5]);
    FE_MUL0(o[7], o[8], f.v[0], f.v[7]);
    asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\tmadc.hi.cc.u32 %1, %6, %7, %1;\n\tmadc.lo.cc.u32 %2, %6, %8, %2;\n\tmadc.hi.cc.u32 %3, %6, %8, %3;\n\tmadc.lo.cc.u32 %4, %6, %9, %4;\n\tmadc.hi.u32 %5, %6, %9, %5;"
        : "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9])
        : "r"(f.v[1]), "r"(f.v[3]), "r"(f.v[5]), "r"(f.v[7]));
    asm("mad.lo.cc.u32 %0, %7, %8, %0;\n\tmadc.hi.cc.u32 %1, %7, %8, %1;\n\tmadc.lo.cc.u32 %2, %7, %9, %2;\n\tmadc.hi.cc.u32 %3, %7, %9, %3;\n\tmadc.lo.cc.u32 %4, %7, %10, %4;\n\tmadc.hi.cc.u32 %5, %7, %10, %5;\n\taddc.u32 %6, 0, 0;"
        : "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9])
        : "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[4]), "r"(f.v[6]));
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\tmadc.hi.cc.u32 %1, %5, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %7, %2;\n\tmadc.hi.cc.u32 %3, %5, %7, %3;\n\taddc.u32 %4, 0, 0;"
        : "+r"(e[6]), "+r"(e[7]), "+r"(e[8]), "+r"(e[9]), "+r"(e[10])
        : "r"(f.v[2]), "r"(f.v[4]), "r"(f.v[6]));
    asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\tmadc.hi.cc.u32 %1, %6, %7, %1;\n\tmadc.lo.cc.u32 %2, %6, %8, %2;\n\tmadc.hi.cc.u32 %3, %6, %8, %3;\n\tmadc.lo.cc.u32 %4, %6, %9, %4;\n\tmadc.hi.u32 %5, %6, %9, %5;"
        : "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10])
        : "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[5]), "r"(f.v[7]));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\tmadc.lo.cc.u32 %2, %4, %6, %2;\n\tmadc.hi.u32 %3, %4, %6, %3;"
        : "+r"(e[8]), "+r"(e[9]), "+r"(e[10]), "+r"(e[11])
        : "r"(f.v[3]), "r"(f.v[5]), "r"(f.v[7]));
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\tmadc.hi.cc.u32 %1, %5, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %7, %2;\n\tmadc.hi.cc.u32 %3, %5, %7, %3;\n\taddc.u32 %4, 0, 0;"
        : "+r"(o[7]), "+r"(o[8]), "+r"(o[9]), "+r"(o[10]), "+r"(o[11])
        : "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[6]));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, 0, 0;"
        : "+r"(e[10]), "+r"(e[11]), "+r"(e[12])
        : "r"(f.v[4]), "r"(f.v[6]));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\tmadc.lo.cc.u32 %2, %4, %6, %2;\n\tmadc.hi.u32 %3, %4, %6, %3;"
        : "+r"(o[9]), "+r"(o[10]), "+r"(o[11]), "+r"(o[12])
        : "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[7]));
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
        : "+r"(e[12]), "+r"(e[13])
        : "r"(f.v[5]), "r"(f.v[7]));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, 0, 0;"
        : "+r"(o[11]), "+r"(o[12]), "+r"(o[13])
        : "r"(f.v[5]), "r"(f.v[6]));
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
        : "+r"(o[13]), "+r"(o[14])
        : "r"(f.v[6]), "r"(f.v[7]));
    t[0] = 0;
    asm("add.cc.u32 %0, %15, %30;\n\taddc.cc.u32 %1, %16, %31;\n\taddc.cc.u32 %2, %17, %32;\n\taddc.cc.u32 %3, %18, %33;\n\taddc.cc.u32 %4, %19, %34;\n\taddc.cc.u32 %5, %20, %35;\n\taddc.cc.u32 %6, %21, %36;\n\taddc.cc.u32 %7, %22, %37;\n\taddc.cc.u32 %8, %23, %38;\n\taddc.cc.u32 %9, %24, %39;\n\taddc.cc.u32 %10, %25, %40;\n\taddc.cc.u32 %11, %26, %41;\n\taddc.cc.u32 %12, %27, %42;\n\taddc.cc.u32 %13, %28, %43;\n\taddc.u32 %14, %29, %44;"
        : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8]), "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11]), "=&r"(t[12]), "=&r"(t[13]), "=&r"(t[14]), "=&r"(t[15])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]), "r"(e[9]), "r"(e[10]), "r"(e[11]), "r"(e[12]), "r"(e[13]), "r"(e[14]), "r"(e[15]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]));
}
#endif
// portable reduction (host)
HD void fe_reduce512_c(fe &h, const uint32_t t[16]) {
    uint64_t c = 0; uint32_t r[8];
    for (int i = 0; i < 8; i++) { c += (uint64_t)t[i] + 38ull * t[8 + i]; r[i] = (uint32_t)c; c >>= 32; }
    c *= 38;
    for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
    r[0] += 38 * (uint32_t)c;
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}

HDMUL void fe_mul(fe &h, const fe &f, const fe &g) {
#if defined(__CUDA_ARCH__)
#if FE_MUL_ROWS
    uint32_t t[16]; fe_mul_rows(t, f, g);
#else
    uint32_t lo[15], hi[15], ex[15];
#pragma unroll
    for (int k = 0; k < 15; k++) {
        const int i0 = k < 8 ? 0 : k - 7, i1 = k < 8 ? k : 7;
        FE_MUL0(lo[k], hi[k], f.v[i0], g.v[k - i0]); ex[k] = 0;
#pragma unroll
        for (int i = i0 + 1; i <= i1; i++) FE_MAC(lo[k], hi[k], ex[k], f.v[i], g.v[k - i]);
    }
    uint32_t t[16]; fe_columns_to_words(t, lo, hi, ex);
#endif
    fe_reduce512(h, t);
#else
    uint32_t t[16]; for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)f.v[i] * g.v[j] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        t[i + 8] = (uint32_t)c;
    }
    fe_reduce512_c(h, t);
#endif
}

HDMUL void fe_sq(fe &h, const fe &f) {
#if defined(__CUDA_ARCH__)
    // off-diagonal products once, doubled by a 1-bit shift of the whole 512-bit value, then the diagonal squares
#if FE_MUL_ROWS
    uint32_t t[16]; fe_sq_rows(t, f);
#else
    uint32_t lo[15], hi[15], ex[15];
    lo[0] = hi[0] = ex[0] = 0; lo[14] = hi[14] = ex[14] = 0;
#pragma unroll
    for (int k = 1; k < 14; k++) {
        const int i0 = k < 8 ? 0 : k - 7;            // pairs i < j = k - i  <=>  i0 <= i < k/2 (+1 if k odd)
        const int i1 = (k - 1) / 2;
        FE_MUL0(lo[k], hi[k], f.v[i0], f.v[k - i0]); ex[k] = 0;
#pragma unroll
        for (int i = i0 + 1; i <= i1; i++) FE_MAC(lo[k], hi[k], ex[k], f.v[i], f.v[k - i]);
    }
    uint32_t t[16]; fe_columns_to_words(t, lo, hi, ex);
#endif
#pragma unroll
    for (int i = 15; i > 0; i--) t[i] = (t[i] << 1) | (t[i - 1] >> 31);
    t[0] <<= 1;
    uint32_t dl[8], dh[8];
#pragma unroll
    for (int i = 0; i < 8; i++) FE_MUL0(dl[i], dh[i], f.v[i], f.v[i]);
    asm("add.cc.u32 %0, %0, %16;\n\taddc.cc.u32 %1, %1, %17;\n\taddc.cc.u32 %2, %2, %18;\n\taddc.cc.u32 %3, %3, %19;\n\taddc.cc.u32 %4, %4, %20;\n\t"
        "addc.cc.u32 %5, %5, %21;\n\taddc.cc.u32 %6, %6, %22;\n\taddc.cc.u32 %7, %7, %23;\n\taddc.cc.u32 %8, %8, %24;\n\taddc.cc.u32 %9, %9, %25;\n\t"
        "addc.cc.u32 %10, %10, %26;\n\taddc.cc.u32 %11, %11, %27;\n\taddc.cc.u32 %12, %12, %28;\n\taddc.cc.u32 %13, %13, %29;\n\taddc.cc.u32 %14, %14, %30;\n\taddc.u32 %15, %15, %31;"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
        : "r"(dl[0]), "r"(dh[0]), "r"(dl[1]), "r"(dh[1]), "r"(dl[2]), "r"(dh[2]), "r"(dl[3]), "r"(dh[3]), "r"(dl[4]), "r"(dh[4]), "r"(dl[5]), "r"(dh[5]), "r"(dl[6]), "r"(dh[6]), "r"(dl[7]), "r"(dh[7]));
    fe_reduce512(h, t);
#else
    fe_mul(h, f, f);
#endif
}
HD void fe_sqn(fe &h, const fe &f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }
// multiply by a small constant (c < 2^26)
HD void fe_mul_small(fe &h, const fe &f, uint32_t c) {
    uint32_t t[16]; uint64_t cy = 0;
    for (int i = 0; i < 8; i++) { cy += (uint64_t)f.v[i] * c; t[i] = (uint32_t)cy; cy >>= 32; }
    t[8] = (uint32_t)cy; for (int i = 9; i < 16; i++) t[i] = 0;
    fe_reduce512_c(h, t);
}

// little-endian bytes, bit 255 ignored (dalek FieldElement::from_bytes)
HD void fe_frombytes(fe &h, const uint8_t *s) {
    for (int i = 0; i < 8; i++) h.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    h.v[7] &= 0x7fffffffu;
}
HD void fe_fromwords(fe &h, const uint32_t w[8]) { for (int i = 0; i < 8; i++) h.v[i] = w[i]; h.v[7] &= 0x7fffffffu; }
// canonical value (< p) as 8 little-endian words
HD void fe_towords(uint32_t w[8], const fe &f) {
    // fold bit 255 (worth 19) twice: value < 2^255 + 19 -> < 2^255 after the second fold unless the value is in [p, 2^255)
    uint32_t r[8]; for (int i = 0; i < 8; i++) r[i] = f.v[i];
    for (int pass = 0; pass < 2; pass++) {
        uint64_t c = 19ull * (r[7] >> 31); r[7] &= 0x7fffffffu;
        for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
    }
    // now r < 2^255; subtract p iff r >= p  <=>  r + 19 >= 2^255
    uint32_t q[8]; uint64_t c = 19;
    for (int i = 0; i < 8; i++) { c += r[i]; q[i] = (uint32_t)c; c >>= 32; }
    const bool ge = (q[7] >> 31) != 0;
    q[7] &= 0x7fffffffu;
    for (int i = 0; i < 8; i++) w[i] = ge ? q[i] : r[i];
}
HD void fe_tobytes(uint8_t *s, const fe &f) {
    uint32_t w[8]; fe_towords(w, f);
    for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)w[i]; s[4 * i + 1] = (uint8_t)(w[i] >> 8); s[4 * i + 2] = (uint8_t)(w[i] >> 16); s[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}
HD bool fe_iszero(const fe &f) { uint32_t w[8]; fe_towords(w, f); uint32_t r = 0; for (int i = 0; i < 8; i++) r |= w[i]; return r == 0; }
HD bool fe_eq(const fe &f, const fe &g) { fe t; fe_sub(t, f, g); return fe_iszero(t); }
HD bool fe_isneg(const fe &f) { uint32_t w[8]; fe_towords(w, f); return w[0] & 1; }
HD void fe_cmov(fe &h, const fe &g, bool b) { for (int i = 0; i < 8; i++) h.v[i] = b ? g.v[i] : h.v[i]; }
HD void fe_abs(fe &h, const fe &f) { fe n; fe_neg(n, f); bool neg = fe_isneg(f); h = f; fe_cmov(h, n, neg); }

// z^(2^252-3)
HDNI void fe_pow22523(fe &out, const fe &z) {
    fe t0, t1, t2;
    fe_sq(t0, z); fe_sqn(t1, t0, 2); fe_mul(t1, z, t1); fe_mul(t0, t0, t1); fe_sq(t0, t0); fe_mul(t0, t1, t0);
    fe_sqn(t1, t0, 5); fe_mul(t0, t1, t0);
    fe_sqn(t1, t0, 10); fe_mul(t1, t1, t0);
    fe_sqn(t2, t1, 20); fe_mul(t1, t2, t1);
    fe_sqn(t1, t1, 10); fe_mul(t0, t1, t0);
    fe_sqn(t1, t0, 50); fe_mul(t1, t1, t0);
    fe_sqn(t2, t1, 100); fe_mul(t1, t2, t1);
    fe_sqn(t1, t1, 50); fe_mul(t0, t1, t0);
    fe_sqn(t0, t0, 2); fe_mul(out, t0, z);
}
HDNI void fe_invert(fe &out, const fe &z) {
    fe t, z3;
    fe_pow22523(t, z); fe_sqn(t, t, 3);
    fe_sq(z3, z); fe_mul(z3, z3, z);
    fe_mul(out, t, z3);
}
