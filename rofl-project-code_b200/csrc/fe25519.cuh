// GF(2^255-19) for sm_100a: radix 2^25.5, 10 x u32 limbs (26,25,26,25,... bits), products accumulated
// in 64-bit registers so every partial product is one IMAD.WIDE.U32 on the FMA pipe.
//
// This is the arithmetic the reference reaches through curve25519-dalek-ng `FieldElement`
// (rofl_crypto/Cargo.toml:13; third-party).  Written from the field definition; all functions are
// __host__ __device__ so tests/hostsim can run the identical code on the CPU against the oracle.
//
// Limb-size discipline (unsigned limbs, "scale" = multiple of the reduced bound 2^26 / 2^25):
//   * fe_mul / fe_sq / fe_carry outputs are REDUCED: even limbs < 2^26, odd limbs < 2^25 + 2^19   (scale 1)
//   * fe_add adds scales; fe_sub(a,b) = a + 2p - b needs scale(b) <= 1 (+slack) and gives scale(a)+2;
//     fe_sub4(a,b) = a + 4p - b needs scale(b) <= 3 and gives scale(a)+4
//   * fe_mul(f,g) needs scale(f)*scale(g) <= 28 and scale(g) <= 3 (19*g_i must fit u32);
//     fe_sq(f) needs scale(f) <= 3 (38*f_i must fit u32)
// Define FE_CHECK_BOUNDS (host builds only) to assert these preconditions at run time.
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

#if defined(FE_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
#include <assert.h>
#define FE_ASSERT(c) assert(c)
#else
#define FE_ASSERT(c) ((void)0)
#endif

#if !defined(__CUDACC__)
// host-only builds (tests/hostsim): the few CUDA vector types the storage helpers use
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define __align__(n) alignas(n)
#endif

struct fe { uint32_t v[10]; };

#define FE_M26 0x3ffffffu
#define FE_M25 0x1ffffffu

HD void fe_0(fe &h) { for (int i = 0; i < 10; i++) h.v[i] = 0; }
HD void fe_1(fe &h) { fe_0(h); h.v[0] = 1; }
HD void fe_add(fe &h, const fe &f, const fe &g) { for (int i = 0; i < 10; i++) h.v[i] = f.v[i] + g.v[i]; }
// 2p limbs: 2*(2^26-19), 2*(2^25-1), 2*(2^26-1), ...
HD void fe_sub(fe &h, const fe &f, const fe &g) {
    FE_ASSERT(g.v[0] <= 0x7ffffdau);
    h.v[0] = f.v[0] + 0x7ffffdau - g.v[0];
    for (int i = 1; i < 10; i++) {
        uint32_t b = (i & 1) ? 0x3fffffeu : 0x7fffffeu;
        FE_ASSERT(g.v[i] <= b);
        h.v[i] = f.v[i] + b - g.v[i];
    }
}
// a + 4p - b
HD void fe_sub4(fe &h, const fe &f, const fe &g) {
    FE_ASSERT(g.v[0] <= 0xfffffb4u);
    h.v[0] = f.v[0] + 0xfffffb4u - g.v[0];
    for (int i = 1; i < 10; i++) {
        uint32_t b = (i & 1) ? 0x7fffffcu : 0xffffffcu;
        FE_ASSERT(g.v[i] <= b);
        h.v[i] = f.v[i] + b - g.v[i];
    }
}
HD void fe_neg(fe &h, const fe &f) { fe z; fe_0(z); fe_sub(h, z, f); }

// carry chain on ten 64-bit columns -> reduced limbs
HD void fe_reduce64(fe &h, uint64_t t[10]) {
    uint64_t c;
    c = t[0] >> 26; t[1] += c; h.v[0] = (uint32_t)t[0] & FE_M26;
    c = t[1] >> 25; t[2] += c; h.v[1] = (uint32_t)t[1] & FE_M25;
    c = t[2] >> 26; t[3] += c; h.v[2] = (uint32_t)t[2] & FE_M26;
    c = t[3] >> 25; t[4] += c; h.v[3] = (uint32_t)t[3] & FE_M25;
    c = t[4] >> 26; t[5] += c; h.v[4] = (uint32_t)t[4] & FE_M26;
    c = t[5] >> 25; t[6] += c; h.v[5] = (uint32_t)t[5] & FE_M25;
    c = t[6] >> 26; t[7] += c; h.v[6] = (uint32_t)t[6] & FE_M26;
    c = t[7] >> 25; t[8] += c; h.v[7] = (uint32_t)t[7] & FE_M25;
    c = t[8] >> 26; t[9] += c; h.v[8] = (uint32_t)t[8] & FE_M26;
    c = t[9] >> 25;            h.v[9] = (uint32_t)t[9] & FE_M25;
    // c < 2^39; wrap 19*c into limb 0 and carry once more into limb 1
    uint64_t w = (uint64_t)h.v[0] + c * 19;
    h.v[0] = (uint32_t)w & FE_M26;
    h.v[1] += (uint32_t)(w >> 26);
}
HD void fe_carry(fe &h, const fe &f) {
    uint64_t t[10];
    for (int i = 0; i < 10; i++) t[i] = f.v[i];
    fe_reduce64(h, t);
}

#define M64(a, b) ((uint64_t)(a) * (uint64_t)(b))

HD void fe_mul(fe &h, const fe &f, const fe &g) {
    const uint32_t f0 = f.v[0], f1 = f.v[1], f2 = f.v[2], f3 = f.v[3], f4 = f.v[4], f5 = f.v[5], f6 = f.v[6], f7 = f.v[7], f8 = f.v[8], f9 = f.v[9];
    const uint32_t g0 = g.v[0], g1 = g.v[1], g2 = g.v[2], g3 = g.v[3], g4 = g.v[4], g5 = g.v[5], g6 = g.v[6], g7 = g.v[7], g8 = g.v[8], g9 = g.v[9];
#if defined(FE_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
    for (int i = 0; i < 10; i++) { FE_ASSERT((uint64_t)g.v[i] * 19 < (1ull << 32)); FE_ASSERT(f.v[i] < (1u << 31)); }
    { uint32_t mf = 0, mg = 0; for (int i = 0; i < 10; i++) { uint32_t a = f.v[i] >> ((i & 1) ? 25 : 26), b = g.v[i] >> ((i & 1) ? 25 : 26); if (a > mf) mf = a; if (b > mg) mg = b; }
      FE_ASSERT((uint64_t)(mf + 1) * (mg + 1) <= 30); }
#endif
    const uint32_t g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4, g5_19 = 19 * g5, g6_19 = 19 * g6, g7_19 = 19 * g7, g8_19 = 19 * g8, g9_19 = 19 * g9;
    const uint32_t f1_2 = 2 * f1, f3_2 = 2 * f3, f5_2 = 2 * f5, f7_2 = 2 * f7, f9_2 = 2 * f9;
    uint64_t t[10];
    t[0] = M64(f0, g0) + M64(f1_2, g9_19) + M64(f2, g8_19) + M64(f3_2, g7_19) + M64(f4, g6_19) + M64(f5_2, g5_19) + M64(f6, g4_19) + M64(f7_2, g3_19) + M64(f8, g2_19) + M64(f9_2, g1_19);
    t[1] = M64(f0, g1) + M64(f1, g0) + M64(f2, g9_19) + M64(f3, g8_19) + M64(f4, g7_19) + M64(f5, g6_19) + M64(f6, g5_19) + M64(f7, g4_19) + M64(f8, g3_19) + M64(f9, g2_19);
    t[2] = M64(f0, g2) + M64(f1_2, g1) + M64(f2, g0) + M64(f3_2, g9_19) + M64(f4, g8_19) + M64(f5_2, g7_19) + M64(f6, g6_19) + M64(f7_2, g5_19) + M64(f8, g4_19) + M64(f9_2, g3_19);
    t[3] = M64(f0, g3) + M64(f1, g2) + M64(f2, g1) + M64(f3, g0) + M64(f4, g9_19) + M64(f5, g8_19) + M64(f6, g7_19) + M64(f7, g6_19) + M64(f8, g5_19) + M64(f9, g4_19);
    t[4] = M64(f0, g4) + M64(f1_2, g3) + M64(f2, g2) + M64(f3_2, g1) + M64(f4, g0) + M64(f5_2, g9_19) + M64(f6, g8_19) + M64(f7_2, g7_19) + M64(f8, g6_19) + M64(f9_2, g5_19);
    t[5] = M64(f0, g5) + M64(f1, g4) + M64(f2, g3) + M64(f3, g2) + M64(f4, g1) + M64(f5, g0) + M64(f6, g9_19) + M64(f7, g8_19) + M64(f8, g7_19) + M64(f9, g6_19);
    t[6] = M64(f0, g6) + M64(f1_2, g5) + M64(f2, g4) + M64(f3_2, g3) + M64(f4, g2) + M64(f5_2, g1) + M64(f6, g0) + M64(f7_2, g9_19) + M64(f8, g8_19) + M64(f9_2, g7_19);
    t[7] = M64(f0, g7) + M64(f1, g6) + M64(f2, g5) + M64(f3, g4) + M64(f4, g3) + M64(f5, g2) + M64(f6, g1) + M64(f7, g0) + M64(f8, g9_19) + M64(f9, g8_19);
    t[8] = M64(f0, g8) + M64(f1_2, g7) + M64(f2, g6) + M64(f3_2, g5) + M64(f4, g4) + M64(f5_2, g3) + M64(f6, g2) + M64(f7_2, g1) + M64(f8, g0) + M64(f9_2, g9_19);
    t[9] = M64(f0, g9) + M64(f1, g8) + M64(f2, g7) + M64(f3, g6) + M64(f4, g5) + M64(f5, g4) + M64(f6, g3) + M64(f7, g2) + M64(f8, g1) + M64(f9, g0);
    fe_reduce64(h, t);
}

HD void fe_sq(fe &h, const fe &f) {
    const uint32_t f0 = f.v[0], f1 = f.v[1], f2 = f.v[2], f3 = f.v[3], f4 = f.v[4], f5 = f.v[5], f6 = f.v[6], f7 = f.v[7], f8 = f.v[8], f9 = f.v[9];
#if defined(FE_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
    for (int i = 0; i < 10; i++) FE_ASSERT(f.v[i] < (3u << ((i & 1) ? 25 : 26)) + (1u << 20));      // scale <= 3
    FE_ASSERT((uint64_t)f5 * 38 < (1ull << 32) && (uint64_t)f7 * 38 < (1ull << 32) && (uint64_t)f9 * 38 < (1ull << 32));
    FE_ASSERT((uint64_t)f6 * 19 < (1ull << 32) && (uint64_t)f8 * 19 < (1ull << 32));
#endif
    const uint32_t f0_2 = 2 * f0, f1_2 = 2 * f1, f2_2 = 2 * f2, f3_2 = 2 * f3, f4_2 = 2 * f4, f5_2 = 2 * f5, f6_2 = 2 * f6, f7_2 = 2 * f7;
    const uint32_t f5_38 = 38 * f5, f6_19 = 19 * f6, f7_38 = 38 * f7, f8_19 = 19 * f8, f9_38 = 38 * f9;
    uint64_t t[10];
    t[0] = M64(f0, f0) + M64(f1_2, f9_38) + M64(f2_2, f8_19) + M64(f3_2, f7_38) + M64(f4_2, f6_19) + M64(f5, f5_38);
    t[1] = M64(f0_2, f1) + M64(f2, f9_38) + M64(f3_2, f8_19) + M64(f4, f7_38) + M64(f5_2, f6_19);
    t[2] = M64(f0_2, f2) + M64(f1_2, f1) + M64(f3_2, f9_38) + M64(f4_2, f8_19) + M64(f5_2, f7_38) + M64(f6, f6_19);
    t[3] = M64(f0_2, f3) + M64(f1_2, f2) + M64(f4, f9_38) + M64(f5_2, f8_19) + M64(f6, f7_38);
    t[4] = M64(f0_2, f4) + M64(f1_2, f3_2) + M64(f2, f2) + M64(f5_2, f9_38) + M64(f6_2, f8_19) + M64(f7, f7_38);
    t[5] = M64(f0_2, f5) + M64(f1_2, f4) + M64(f2_2, f3) + M64(f6, f9_38) + M64(f7_2, f8_19);
    t[6] = M64(f0_2, f6) + M64(f1_2, f5_2) + M64(f2_2, f4) + M64(f3_2, f3) + M64(f7_2, f9_38) + M64(f8, f8_19);
    t[7] = M64(f0_2, f7) + M64(f1_2, f6) + M64(f2_2, f5) + M64(f3_2, f4) + M64(f8, f9_38);
    t[8] = M64(f0_2, f8) + M64(f1_2, f7_2) + M64(f2_2, f6) + M64(f3_2, f5_2) + M64(f4, f4) + M64(f9, f9_38);
    t[9] = M64(f0_2, f9) + M64(f1_2, f8) + M64(f2_2, f7) + M64(f3_2, f6) + M64(f4_2, f5);
    fe_reduce64(h, t);
}
HD void fe_sqn(fe &h, const fe &f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }
// small-constant multiply (c*scale must stay in range); result reduced
HD void fe_mul_small(fe &h, const fe &f, uint32_t c) {
    uint64_t t[10];
    for (int i = 0; i < 10; i++) t[i] = M64(f.v[i], c);
    fe_reduce64(h, t);
}

// little-endian bytes, bit 255 ignored (dalek FieldElement::from_bytes)
HD void fe_frombytes(fe &h, const uint8_t *s) {
    uint32_t w[8];
    for (int i = 0; i < 8; i++) w[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    // limb offsets (bits): 0,26,51,77,102,128,153,179,204,230
    h.v[0] = w[0] & FE_M26;
    h.v[1] = ((w[0] >> 26) | (w[1] << 6)) & FE_M25;
    h.v[2] = ((w[1] >> 19) | (w[2] << 13)) & FE_M26;
    h.v[3] = ((w[2] >> 13) | (w[3] << 19)) & FE_M25;
    h.v[4] = (w[3] >> 6) & FE_M26;
    h.v[5] = w[4] & FE_M25;
    h.v[6] = ((w[4] >> 25) | (w[5] << 7)) & FE_M26;
    h.v[7] = ((w[5] >> 19) | (w[6] << 13)) & FE_M25;
    h.v[8] = ((w[6] >> 12) | (w[7] << 20)) & FE_M26;
    h.v[9] = (w[7] >> 6) & FE_M25;
}
// from 8 little-endian 32-bit words
HD void fe_fromwords(fe &h, const uint32_t w[8]) {
    h.v[0] = w[0] & FE_M26;
    h.v[1] = ((w[0] >> 26) | (w[1] << 6)) & FE_M25;
    h.v[2] = ((w[1] >> 19) | (w[2] << 13)) & FE_M26;
    h.v[3] = ((w[2] >> 13) | (w[3] << 19)) & FE_M25;
    h.v[4] = (w[3] >> 6) & FE_M26;
    h.v[5] = w[4] & FE_M25;
    h.v[6] = ((w[4] >> 25) | (w[5] << 7)) & FE_M26;
    h.v[7] = ((w[5] >> 19) | (w[6] << 13)) & FE_M25;
    h.v[8] = ((w[6] >> 12) | (w[7] << 20)) & FE_M26;
    h.v[9] = (w[7] >> 6) & FE_M25;
}
// canonical value as 8 little-endian words
HD void fe_towords(uint32_t w[8], const fe &f) {
    fe t; fe_carry(t, f); fe_carry(t, t);          // limbs now tight: even < 2^26, odd < 2^25 (+1 on limb 1)
    // q = 1 iff t >= p : propagate t + 19 through the limbs
    uint32_t q = (t.v[0] + 19) >> 26;
    q = (t.v[1] + q) >> 25; q = (t.v[2] + q) >> 26; q = (t.v[3] + q) >> 25; q = (t.v[4] + q) >> 26;
    q = (t.v[5] + q) >> 25; q = (t.v[6] + q) >> 26; q = (t.v[7] + q) >> 25; q = (t.v[8] + q) >> 26; q = (t.v[9] + q) >> 25;
    t.v[0] += 19 * q;
    uint32_t c;
    c = t.v[0] >> 26; t.v[0] &= FE_M26; t.v[1] += c;
    c = t.v[1] >> 25; t.v[1] &= FE_M25; t.v[2] += c;
    c = t.v[2] >> 26; t.v[2] &= FE_M26; t.v[3] += c;
    c = t.v[3] >> 25; t.v[3] &= FE_M25; t.v[4] += c;
    c = t.v[4] >> 26; t.v[4] &= FE_M26; t.v[5] += c;
    c = t.v[5] >> 25; t.v[5] &= FE_M25; t.v[6] += c;
    c = t.v[6] >> 26; t.v[6] &= FE_M26; t.v[7] += c;
    c = t.v[7] >> 25; t.v[7] &= FE_M25; t.v[8] += c;
    c = t.v[8] >> 26; t.v[8] &= FE_M26; t.v[9] += c;
    t.v[9] &= FE_M25;
    w[0] = t.v[0] | (t.v[1] << 26);
    w[1] = (t.v[1] >> 6) | (t.v[2] << 19);
    w[2] = (t.v[2] >> 13) | (t.v[3] << 13);
    w[3] = (t.v[3] >> 19) | (t.v[4] << 6);
    w[4] = t.v[5] | (t.v[6] << 25);
    w[5] = (t.v[6] >> 7) | (t.v[7] << 19);
    w[6] = (t.v[7] >> 13) | (t.v[8] << 12);
    w[7] = (t.v[8] >> 20) | (t.v[9] << 6);
}
HD void fe_tobytes(uint8_t *s, const fe &f) {
    uint32_t w[8]; fe_towords(w, f);
    for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)w[i]; s[4 * i + 1] = (uint8_t)(w[i] >> 8); s[4 * i + 2] = (uint8_t)(w[i] >> 16); s[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}
HD bool fe_iszero(const fe &f) { uint32_t w[8]; fe_towords(w, f); uint32_t r = 0; for (int i = 0; i < 8; i++) r |= w[i]; return r == 0; }
HD bool fe_eq(const fe &f, const fe &g) { fe c, t; fe_carry(c, g); fe_sub(t, f, c); return fe_iszero(t); }
HD bool fe_isneg(const fe &f) { uint32_t w[8]; fe_towords(w, f); return w[0] & 1; }
HD void fe_cmov(fe &h, const fe &g, bool b) { for (int i = 0; i < 10; i++) h.v[i] = b ? g.v[i] : h.v[i]; }
// input must be reduced
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
