// Integers mod l = 2^252 + 27742317777372353535851937790883648493 (ristretto255 group order) on 8 x u32 limbs.
// curve25519-dalek-ng `Scalar` semantics (32-byte little-endian canonical form at every API edge).
// Multiplication is two Montgomery REDC passes (R = 2^256) so values stay canonical between calls;
// the oracle uses a different method (2^252 = -c folding on 64-bit limbs), so the two cross-check.
// All __host__ __device__ (see fe25519.cuh).
#pragma once
#include "fe25519.cuh"
#include "constants32.cuh"

struct sc { uint32_t v[8]; };

#if defined(__CUDA_ARCH__)
#define SC_CONST static __device__ __constant__ const
#else
#define SC_CONST static const
#endif
SC_CONST uint32_t SC_L_[8] = SC_L;
SC_CONST uint32_t SC_R2_[8] = SC_MONT_R2;
#define SC_NINV SC_MONT_NINV

HD void sc_0(sc &r) { for (int i = 0; i < 8; i++) r.v[i] = 0; }
HD void sc_from_u64(sc &r, uint64_t x) { sc_0(r); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); }
HD bool sc_iszero(const sc &a) { uint32_t r = 0; for (int i = 0; i < 8; i++) r |= a.v[i]; return r == 0; }
HD bool sc_eq(const sc &a, const sc &b) { uint32_t r = 0; for (int i = 0; i < 8; i++) r |= a.v[i] ^ b.v[i]; return r == 0; }
HD void sc_frombytes(sc &r, const uint8_t *s) {
    for (int i = 0; i < 8; i++) r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
}
HD void sc_tobytes(uint8_t *s, const sc &a) {
    for (int i = 0; i < 8; i++) { s[4 * i] = (uint8_t)a.v[i]; s[4 * i + 1] = (uint8_t)(a.v[i] >> 8); s[4 * i + 2] = (uint8_t)(a.v[i] >> 16); s[4 * i + 3] = (uint8_t)(a.v[i] >> 24); }
}
// a >= l ?
HD bool sc_geq_l(const uint32_t a[8]) {
    for (int i = 7; i >= 0; i--) { if (a[i] > SC_L_[i]) return true; if (a[i] < SC_L_[i]) return false; }
    return true;
}
HD bool sc_is_canonical(const sc &a) { return !sc_geq_l(a.v); }
// r = a - l if a >= l  (a < 2l)
HD void sc_cond_sub_l(uint32_t a[8], uint32_t extra_hi = 0) {
    uint32_t t[8]; uint64_t br = 0;
    for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a[i] - SC_L_[i] - br; t[i] = (uint32_t)d; br = (d >> 32) & 1; }
    bool ge = (extra_hi != 0) || (br == 0);
    for (int i = 0; i < 8; i++) a[i] = ge ? t[i] : a[i];
}
HD void sc_add(sc &r, const sc &a, const sc &b) {
    uint64_t c = 0; uint32_t t[8];
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    sc_cond_sub_l(t);                              // a + b < 2l < 2^254: no carry out
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
}
HD void sc_sub(sc &r, const sc &a, const sc &b) {
    uint64_t br = 0; uint32_t t[8];
    for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)a.v[i] - b.v[i] - br; t[i] = (uint32_t)d; br = (d >> 32) & 1; }
    uint64_t c = 0; uint32_t m = br ? 0xffffffffu : 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)t[i] + (SC_L_[i] & m); r.v[i] = (uint32_t)c; c >>= 32; }
}
HD void sc_neg(sc &r, const sc &a) { sc z; sc_0(z); sc_sub(r, z, a); }
// Montgomery product a*b/2^256 mod l (CIOS); needs a*b < l*2^256; result < l
HDNI void sc_montmul(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t t[10];
    for (int i = 0; i < 10; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)a[i] * b[j] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * SC_NINV;
        c = (uint64_t)m * SC_L_[0] + t[0]; c >>= 32;
        for (int j = 1; j < 8; j++) { c += (uint64_t)m * SC_L_[j] + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32);
    }
    sc_cond_sub_l(t, t[8]);
    for (int i = 0; i < 8; i++) r[i] = t[i];
}
#if defined(__CUDA_ARCH__)
// Device: the 512-bit product by the field's row-wise IMAD.WIDE routine, then l = 2^252 + c (c < 2^125) folded twice:
//   x = xl + 2^252 xh  ==  xl - c xh;   c xh = p1l + 2^252 p1h  ==  p1l - c p1h   =>   x == xl - p1l + c p1h  in (-2^252, 2^253)
// 64 + 32 + 16 multiplications instead of the 256 of two Montgomery passes (measured, one lane: 3683 -> ~900 cycles).
__device__ __forceinline__ void sc_mul_small(uint32_t *out, const uint32_t *cw, const uint32_t *x, int nx) {      // out[0 .. nx+4) = c[0..4) * x[0..nx)
    for (int i = 0; i < nx + 4; i++) out[i] = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t carry = 0;
        for (int j = 0; j < nx; j++) {
            const uint64_t acc = (uint64_t)cw[i] * x[j] + out[i + j] + carry;
            out[i + j] = (uint32_t)acc; carry = (uint32_t)(acc >> 32);
        }
        out[i + nx] = carry;
    }
}
__device__ __forceinline__ void sc_mul_dev(sc &r, const sc &a, const sc &b) {        // any 256-bit a, b
    uint32_t t[16];
    { fe fa, fb; for (int i = 0; i < 8; i++) { fa.v[i] = a.v[i]; fb.v[i] = b.v[i]; } fe_mul_rows(t, fa, fb); }
    uint32_t cw[4]; for (int i = 0; i < 4; i++) cw[i] = SC_L_[i];
    uint32_t xh[9], p1[13], p1h[5], p2[9];
    for (int i = 0; i < 8; i++) xh[i] = (t[7 + i] >> 28) | (t[8 + i] << 4);
    xh[8] = t[15] >> 28;
    sc_mul_small(p1, cw, xh, 9);                                  // < 2^385
    for (int i = 0; i < 5; i++) p1h[i] = (p1[7 + i] >> 28) | (7 + i + 1 < 13 ? p1[8 + i] << 4 : 0);
    sc_mul_small(p2, cw, p1h, 5);                                 // < 2^258
    const uint32_t p2h = (p2[7] >> 28) | (p2[8] << 4);            // < 2^6:  p2 == p2l - c p2h
    int64_t acc = 0; uint32_t o[8];
    uint64_t cp = 0;                                              // running c * p2h
    for (int i = 0; i < 8; i++) {
        const uint32_t xl = i < 7 ? t[i] : (t[7] & 0x0fffffffu), pl = i < 7 ? p1[i] : (p1[7] & 0x0fffffffu), ql = i < 7 ? p2[i] : (p2[7] & 0x0fffffffu);
        if (i < 4) cp += (uint64_t)cw[i] * p2h;
        acc += (int64_t)xl - (int64_t)pl + (int64_t)ql - (int64_t)(uint32_t)cp;
        cp >>= 32;
        o[i] = (uint32_t)acc; acc >>= 32;
    }
    for (int k = 0; k < 2 && acc < 0; k++) { uint64_t c = 0; for (int i = 0; i < 8; i++) { c += (uint64_t)o[i] + SC_L_[i]; o[i] = (uint32_t)c; c >>= 32; } acc += (int64_t)c; }
    sc_cond_sub_l(o);
    for (int i = 0; i < 8; i++) r.v[i] = o[i];
}
#endif
HD void sc_mul(sc &r, const sc &a, const sc &b) {
#if defined(__CUDA_ARCH__)
    sc_mul_dev(r, a, b);
#else
    uint32_t t[8];
    sc_montmul(t, a.v, b.v);          // a b / R
    sc_montmul(r.v, t, SC_R2_);       // a b
#endif
}
HD void sc_muladd(sc &r, const sc &a, const sc &b, const sc &c) { sc t; sc_mul(t, a, b); sc_add(r, t, c); }
HD void sc_sq(sc &r, const sc &a) { sc_mul(r, a, a); }
// reduce a 512-bit little-endian value (16 words): lo + hi*R == mont(lo, R mod l ... ) -- done as
// mont(lo, R2)*1 path: x mod l = mont(mont(lo,R2),1)?  Simpler: lo mod l + mont(hi, R2) (= hi*R mod l)
HD void sc_from_wide_words(sc &r, const uint32_t w[16]) {
    uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    uint32_t lo_m[8], lo[8], hi[8];
    sc_montmul(lo_m, w, SC_R2_);      // lo * R   (lo < 2^256, R2 < l  => product < l*2^256)
    sc_montmul(lo, lo_m, one);        // lo mod l
    sc_montmul(hi, w + 8, SC_R2_);    // hi * R mod l
    sc a, b; for (int i = 0; i < 8; i++) { a.v[i] = lo[i]; b.v[i] = hi[i]; }
    sc_add(r, a, b);
}
HD void sc_from_bytes_wide(sc &r, const uint8_t *s) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    sc_from_wide_words(r, w);
}
// reduce an arbitrary 256-bit value
HD void sc_from_bytes_mod_order(sc &r, const uint8_t *s) {
    uint32_t w[16]; for (int i = 0; i < 16; i++) w[i] = 0;
    for (int i = 0; i < 8; i++) w[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    sc_from_wide_words(r, w);
}
// a^(l-2)
HD void sc_invert(sc &r, const sc &a) {
    uint32_t e[8]; for (int i = 0; i < 8; i++) e[i] = SC_L_[i];
    e[0] -= 2;
    sc acc; sc_from_u64(acc, 1);
    for (int i = 252; i >= 0; i--) {
        sc_mul(acc, acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) sc_mul(acc, acc, a);
    }
    r = acc;
}
// a^-1 by the binary extended Euclid (variable time: only used on public Fiat-Shamir challenges); a canonical, non-zero.
// ~750 shift / add / subtract steps on 8-word integers instead of the ~380 modular multiplications of sc_invert.
HD void sc_raw_shr1(uint32_t a[8]) { for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31); a[7] >>= 1; }
HD void sc_raw_add_l(uint32_t a[8]) { uint64_t c = 0; for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + SC_L_[i]; a[i] = (uint32_t)c; c >>= 32; } }
HD bool sc_raw_sub(uint32_t a[8], const uint32_t b[8]) { int64_t bw = 0; for (int i = 0; i < 8; i++) { bw += (int64_t)a[i] - (int64_t)b[i]; a[i] = (uint32_t)bw; bw >>= 32; } return bw != 0; }
HD bool sc_raw_is_one(const uint32_t a[8]) { uint32_t r = a[0] ^ 1; for (int i = 1; i < 8; i++) r |= a[i]; return r == 0; }
HD void sc_invert_euclid(sc &r, const sc &a) {
    if (sc_iszero(a)) { sc_0(r); return; }
    uint32_t u[8], v[8], x1[8], x2[8];
    for (int i = 0; i < 8; i++) { u[i] = a.v[i]; v[i] = SC_L_[i]; x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    for (;;) {
        if (sc_raw_is_one(u)) { for (int i = 0; i < 8; i++) r.v[i] = x1[i]; return; }
        if (sc_raw_is_one(v)) { for (int i = 0; i < 8; i++) r.v[i] = x2[i]; return; }
        while (!(u[0] & 1)) { sc_raw_shr1(u); if (x1[0] & 1) sc_raw_add_l(x1); sc_raw_shr1(x1); }       // x1 + l < 2^254
        while (!(v[0] & 1)) { sc_raw_shr1(v); if (x2[0] & 1) sc_raw_add_l(x2); sc_raw_shr1(x2); }
        bool ge = true;                                    // u >= v ?
        for (int i = 7; i >= 0; i--) { if (u[i] != v[i]) { ge = u[i] > v[i]; break; } }
        if (ge) { sc_raw_sub(u, v); if (sc_raw_sub(x1, x2)) sc_raw_add_l(x1); }
        else { sc_raw_sub(v, u); if (sc_raw_sub(x2, x1)) sc_raw_add_l(x2); }
    }
}
// a^-1 by Bernstein-Yang division steps ("safegcd"), variable time -- used on public Fiat-Shamir challenges only.  The binary Euclid above
// walks ~750 dependent steps over 8-word integers (90 us for one lane on B200, and every IPP round waits for one inversion).  Here 30 division
// steps at a time run on the low 32 bits of (f, g) alone -- several per iteration: a run of zeros is shifted out at once and up to six bits of g
// are cancelled by one multiple of f -- and are applied to the 9 x 30-bit signed-limb integers f, g, d, e as one 2x2 matrix of 32-bit entries, so
// that every product is a single 32 x 32 -> 64 multiply-add (d a = f, e a = g mod l throughout; f ends as +-1).  At most 25 batches for 256-bit
// inputs ((49 * 256 + 57) / 17 = 741 steps).  The first version (62-bit limbs, one step per iteration, 128-bit products) took 49 us on one lane.
struct sc_s30 { int32_t v[9]; };                    // sum v[i] 2^(30 i); v[0..7] in [0, 2^30), v[8] signed
#define SC_M30 0x3fffffff
HD void sc_to_s30(sc_s30 &o, const uint32_t w[8]) {
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, wi = bit >> 5, sh = bit & 31;
        uint64_t x = (uint64_t)w[wi] >> sh;
        if (wi + 1 < 8) x |= (uint64_t)w[wi + 1] << (32 - sh);
        o.v[i] = (int32_t)((uint32_t)x & (uint32_t)SC_M30);
    }
}
HD void sc_s30_add(sc_s30 &d, const sc_s30 &m, int32_t sign) {       // d += sign * m  (sign = +-1)
    int32_t c = 0;
    for (int i = 0; i < 8; i++) { const int32_t t = d.v[i] + sign * m.v[i] + c; d.v[i] = t & SC_M30; c = t >> 30; }
    d.v[8] += sign * m.v[8] + c;
}
HD int sc_ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// t = {u, v, q, r} with 2^30 f' = u f + v g, 2^30 g' = q f + r g after 30 division steps on the low words; returns the new delta
HD int32_t sc_divsteps30(int32_t delta, uint32_t f, uint32_t g, int32_t t[4]) {
    int32_t u = 1, v = 0, q = 0, r = 1; int i = 30;
    for (;;) {
        const int z = sc_ctz32(g | (~0u << i));       // the even steps: (delta, f, g) -> (1 + delta, f, g / 2)
        g >>= z; u = (int32_t)((uint32_t)u << z); v = (int32_t)((uint32_t)v << z); delta += z; i -= z;
        if (i == 0) break;
        if (delta > 0) {                              // g odd, delta > 0: (delta, f, g) -> (1 - delta, g, (g - f) / 2); the halving is counted above
            const uint32_t tf = f; f = g; g = 0u - tf;
            const int32_t tu = u, tv = v; u = q; v = r; q = -tu; r = -tv;
            delta = -delta;
        }
        // delta <= 0: the next 1 - delta steps cannot swap; cancel up to 6 bits of g with w = -g / f mod 2^k  (f^-1 = f (2 - f^2) mod 64)
        int lim = 1 - delta; if (lim > i) lim = i; if (lim > 6) lim = 6;
        const uint32_t w = (f * g * (f * f - 2u)) & ((1u << lim) - 1u);
        g += f * w; q += u * (int32_t)w; r += v * (int32_t)w;
    }
    t[0] = u; t[1] = v; t[2] = q; t[3] = r;
    return delta;
}
HD void sc_invert_vartime(sc &r, const sc &a) {
    if (sc_iszero(a)) { sc_0(r); return; }
    sc_s30 L, f, g, d, e;
    { uint32_t lw[8]; for (int i = 0; i < 8; i++) lw[i] = SC_L_[i]; sc_to_s30(L, lw); }
    f = L; sc_to_s30(g, a.v);
    for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; }
    e.v[0] = 1;
    uint32_t linv = (uint32_t)L.v[0];                 // l^-1 mod 2^32 by Newton (l odd), used mod 2^30
    for (int i = 0; i < 5; i++) linv *= 2u - (uint32_t)L.v[0] * linv;
    int32_t delta = 1;
    for (int it = 0; it < 26; it++) {
        int32_t t[4];
        delta = sc_divsteps30(delta, (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30), t);
        const int64_t u = t[0], v = t[1], q = t[2], rr = t[3];
        {   // (d, e) <- (u d + v e, q d + r e) / 2^30 mod l: a multiple of l makes the low 30 bits vanish
            int64_t cd = u * d.v[0] + v * e.v[0], ce = q * d.v[0] + rr * e.v[0];
            const int64_t kd = (int64_t)(((0u - (uint32_t)cd) * linv) & (uint32_t)SC_M30), ke = (int64_t)(((0u - (uint32_t)ce) * linv) & (uint32_t)SC_M30);
            cd += kd * L.v[0]; ce += ke * L.v[0];
            cd >>= 30; ce >>= 30;
            for (int i = 1; i < 9; i++) {
                cd += u * d.v[i] + v * e.v[i] + kd * L.v[i];
                ce += q * d.v[i] + rr * e.v[i] + ke * L.v[i];
                if (i < 8) { d.v[i - 1] = (int32_t)(cd & SC_M30); e.v[i - 1] = (int32_t)(ce & SC_M30); cd >>= 30; ce >>= 30; }
                else { d.v[7] = (int32_t)(cd & SC_M30); e.v[7] = (int32_t)(ce & SC_M30); d.v[8] = (int32_t)(cd >> 30); e.v[8] = (int32_t)(ce >> 30); }
            }
            if (d.v[8] >= 0) sc_s30_add(d, L, -1);   // keep |d|, |e| < 2 l
            if (e.v[8] >= 0) sc_s30_add(e, L, -1);
        }
        {   // (f, g) <- (u f + v g, q f + r g) / 2^30, exact
            int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + rr * g.v[0];
            cf >>= 30; cg >>= 30;
            for (int i = 1; i < 9; i++) {
                cf += u * f.v[i] + v * g.v[i]; cg += q * f.v[i] + rr * g.v[i];
                if (i < 8) { f.v[i - 1] = (int32_t)(cf & SC_M30); g.v[i - 1] = (int32_t)(cg & SC_M30); cf >>= 30; cg >>= 30; }
                else { f.v[7] = (int32_t)(cf & SC_M30); g.v[7] = (int32_t)(cg & SC_M30); f.v[8] = (int32_t)(cf >> 30); g.v[8] = (int32_t)(cg >> 30); }
            }
        }
        int32_t nz = 0; for (int i = 0; i < 9; i++) nz |= g.v[i];
        if (nz == 0) break;
    }
    if (f.v[8] < 0) { sc_s30 z; for (int i = 0; i < 9; i++) { z.v[i] = d.v[i]; d.v[i] = 0; } sc_s30_add(d, z, -1); }       // f = -1: negate
    for (int k = 0; k < 3 && d.v[8] < 0; k++) sc_s30_add(d, L, 1);
    for (int k = 0; k < 3; k++) { sc_s30 t2 = d; sc_s30_add(t2, L, -1); if (t2.v[8] < 0) break; d = t2; }
    uint64_t acc = 0; int bits = 0, wi = 0;
    for (int i = 0; i < 9; i++) {
        acc |= (uint64_t)(uint32_t)d.v[i] << bits; bits += 30;
        while (bits >= 32 && wi < 8) { r.v[wi++] = (uint32_t)acc; acc >>= 32; bits -= 32; }
    }
    if (wi < 8) r.v[wi] = (uint32_t)acc;
}
HD void sc_pow_u64(sc &r, const sc &a, uint64_t e) {
    sc acc, base = a; sc_from_u64(acc, 1);
    while (e) { if (e & 1) sc_mul(acc, acc, base); sc_mul(base, base, base); e >>= 1; }
    r = acc;
}
