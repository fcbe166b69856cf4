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
// walks ~750 dependent steps over 8-word integers (90 us for one thread on B200, and every IPP round waits for one inversion); here 62 division
// steps at a time run on the low 64 bits of (f, g) alone and are applied to the 5 x 62-bit signed-limb integers f, g, d, e as one 2x2 matrix
// (d a = f, e a = g mod l throughout; f ends as +-1).  At most 12 batches for 256-bit inputs ((49 * 256 + 57) / 17 = 741 steps).
typedef __int128 sc_i128;
struct sc_s62 { int64_t v[5]; };                    // sum v[i] 2^(62 i); v[0..3] in [0, 2^62), v[4] signed
#define SC_M62 ((int64_t)((1ULL << 62) - 1))
HD void sc_to_s62(sc_s62 &o, const uint32_t w[8]) {
    uint64_t q[4]; for (int i = 0; i < 4; i++) q[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    o.v[0] = (int64_t)(q[0] & SC_M62);
    o.v[1] = (int64_t)(((q[0] >> 62) | (q[1] << 2)) & SC_M62);
    o.v[2] = (int64_t)(((q[1] >> 60) | (q[2] << 4)) & SC_M62);
    o.v[3] = (int64_t)(((q[2] >> 58) | (q[3] << 6)) & SC_M62);
    o.v[4] = (int64_t)(q[3] >> 56);
}
HD void sc_s62_add(sc_s62 &d, const sc_s62 &m, int64_t sign) {       // d += sign * m  (sign = +-1)
    int64_t c = 0;
    for (int i = 0; i < 4; i++) { const int64_t t = d.v[i] + sign * m.v[i] + c; d.v[i] = t & SC_M62; c = t >> 62; }
    d.v[4] += sign * m.v[4] + c;
}
// t = {u, v, q, r} with 2^62 f' = u f + v g, 2^62 g' = q f + r g after 62 division steps on the low words; returns the new delta
HD int64_t sc_divsteps62(int64_t delta, uint64_t f, uint64_t g, int64_t t[4]) {
    int64_t u = 1, v = 0, q = 0, r = 1;
    for (int i = 0; i < 62; i++) {
        if (g & 1) {
            if (delta > 0) {                          // (delta, f, g) -> (1 - delta, g, (g - f) / 2)
                const uint64_t nf = g; g = (g - f) >> 1; f = nf;
                const int64_t nq = q - u, nr = r - v; u = 2 * q; v = 2 * r; q = nq; r = nr;
                delta = 1 - delta;
            } else {                                  // (1 + delta, f, (g + f) / 2)
                g = (g + f) >> 1; q += u; r += v; u *= 2; v *= 2; delta += 1;
            }
        } else { g >>= 1; u *= 2; v *= 2; delta += 1; }
    }
    t[0] = u; t[1] = v; t[2] = q; t[3] = r;
    return delta;
}
HD void sc_invert_vartime(sc &r, const sc &a) {
    if (sc_iszero(a)) { sc_0(r); return; }
    sc_s62 L, f, g, d, e;
    { uint32_t lw[8]; for (int i = 0; i < 8; i++) lw[i] = SC_L_[i]; sc_to_s62(L, lw); }
    f = L; sc_to_s62(g, a.v);
    for (int i = 0; i < 5; i++) { d.v[i] = 0; e.v[i] = 0; }
    e.v[0] = 1;
    uint64_t linv = (uint64_t)L.v[0];                 // l^-1 mod 2^64 by Newton (l odd), used mod 2^62
    for (int i = 0; i < 6; i++) linv *= 2 - (uint64_t)L.v[0] * linv;
    int64_t delta = 1;
    for (int it = 0; it < 13; it++) {
        int64_t t[4];
        delta = sc_divsteps62(delta, (uint64_t)f.v[0] | ((uint64_t)f.v[1] << 62), (uint64_t)g.v[0] | ((uint64_t)g.v[1] << 62), t);
        const int64_t u = t[0], v = t[1], q = t[2], rr = t[3];
        {   // (d, e) <- (u d + v e, q d + r e) / 2^62 mod l: a multiple of l makes the low 62 bits vanish
            sc_i128 cd = (sc_i128)u * d.v[0] + (sc_i128)v * e.v[0], ce = (sc_i128)q * d.v[0] + (sc_i128)rr * e.v[0];
            const int64_t kd = (int64_t)((0 - (uint64_t)cd * linv) & (uint64_t)SC_M62), ke = (int64_t)((0 - (uint64_t)ce * linv) & (uint64_t)SC_M62);
            cd += (sc_i128)kd * L.v[0]; ce += (sc_i128)ke * L.v[0];
            cd >>= 62; ce >>= 62;
            for (int i = 1; i < 5; i++) {
                cd += (sc_i128)u * d.v[i] + (sc_i128)v * e.v[i] + (sc_i128)kd * L.v[i];
                ce += (sc_i128)q * d.v[i] + (sc_i128)rr * e.v[i] + (sc_i128)ke * L.v[i];
                if (i < 4) { d.v[i - 1] = (int64_t)cd & SC_M62; e.v[i - 1] = (int64_t)ce & SC_M62; cd >>= 62; ce >>= 62; }
                else {
                    const int64_t d3 = (int64_t)cd & SC_M62, e3 = (int64_t)ce & SC_M62; cd >>= 62; ce >>= 62;
                    // (d.v[i-1] must stay readable until both sums have used d.v[i], hence the late stores)
                    d.v[3] = d3; e.v[3] = e3; d.v[4] = (int64_t)cd; e.v[4] = (int64_t)ce;
                }
            }
            if (d.v[4] >= 0) sc_s62_add(d, L, -1);   // keep |d|, |e| < 2 l
            if (e.v[4] >= 0) sc_s62_add(e, L, -1);
        }
        {   // (f, g) <- (u f + v g, q f + r g) / 2^62, exact
            sc_i128 cf = (sc_i128)u * f.v[0] + (sc_i128)v * g.v[0], cg = (sc_i128)q * f.v[0] + (sc_i128)rr * g.v[0];
            cf >>= 62; cg >>= 62;
            for (int i = 1; i < 5; i++) {
                cf += (sc_i128)u * f.v[i] + (sc_i128)v * g.v[i]; cg += (sc_i128)q * f.v[i] + (sc_i128)rr * g.v[i];
                if (i < 4) { f.v[i - 1] = (int64_t)cf & SC_M62; g.v[i - 1] = (int64_t)cg & SC_M62; cf >>= 62; cg >>= 62; }
                else { const int64_t f3 = (int64_t)cf & SC_M62, g3 = (int64_t)cg & SC_M62; cf >>= 62; cg >>= 62; f.v[3] = f3; g.v[3] = g3; f.v[4] = (int64_t)cf; g.v[4] = (int64_t)cg; }
            }
        }
        if ((g.v[0] | g.v[1] | g.v[2] | g.v[3] | g.v[4]) == 0) break;
    }
    if (f.v[4] < 0) { sc_s62 z; for (int i = 0; i < 5; i++) { z.v[i] = d.v[i]; d.v[i] = 0; } sc_s62_add(d, z, -1); }       // f = -1: negate
    for (int k = 0; k < 3 && d.v[4] < 0; k++) sc_s62_add(d, L, 1);
    for (int k = 0; k < 3; k++) { sc_s62 t2 = d; sc_s62_add(t2, L, -1); if (t2.v[4] < 0) break; d = t2; }
    const uint64_t q0 = (uint64_t)d.v[0] | ((uint64_t)d.v[1] << 62), q1 = ((uint64_t)d.v[1] >> 2) | ((uint64_t)d.v[2] << 60),
                   q2 = ((uint64_t)d.v[2] >> 4) | ((uint64_t)d.v[3] << 58), q3 = ((uint64_t)d.v[3] >> 6) | ((uint64_t)d.v[4] << 56);
    r.v[0] = (uint32_t)q0; r.v[1] = (uint32_t)(q0 >> 32); r.v[2] = (uint32_t)q1; r.v[3] = (uint32_t)(q1 >> 32);
    r.v[4] = (uint32_t)q2; r.v[5] = (uint32_t)(q2 >> 32); r.v[6] = (uint32_t)q3; r.v[7] = (uint32_t)(q3 >> 32);
}
HD void sc_pow_u64(sc &r, const sc &a, uint64_t e) {
    sc acc, base = a; sc_from_u64(acc, 1);
    while (e) { if (e & 1) sc_mul(acc, acc, base); sc_mul(base, base, base); e >>= 1; }
    r = acc;
}
