/* ORACLE -- TEST INFRASTRUCTURE ONLY. Never linked into, imported by or called from the product
 * (rofl-project-code_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker or the reported CPU baseline.
 *
 * GF(2^255-19) in radix 2^51 (5 x u64, products in unsigned __int128).  This is the arithmetic
 * the reference reaches through curve25519-dalek-ng 4.1.1 `FieldElement` (third-party crate, NOT
 * under /root/reference; Cargo.lock:363-364).  Restated from the field's definition; the
 * representation is deliberately different from the device code (radix 2^25.5) so the two
 * implementations cross-check each other.  Pinned against libsodium in tests/test_oracle_*.py.
 */
#ifndef ROFL_ORACLE_FE51_H
#define ROFL_ORACLE_FE51_H
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t v[5]; } fe;
typedef unsigned __int128 u128;

#define FE_MASK51 ((1ULL << 51) - 1)

static inline void fe_0(fe *h) { memset(h, 0, sizeof *h); }
static inline void fe_1(fe *h) { fe_0(h); h->v[0] = 1; }
static inline void fe_copy(fe *h, const fe *f) { *h = *f; }

static inline void fe_carry(fe *h) {
    uint64_t c;
    c = h->v[0] >> 51; h->v[0] &= FE_MASK51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= FE_MASK51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= FE_MASK51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= FE_MASK51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= FE_MASK51; h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= FE_MASK51; h->v[1] += c;
}
static inline void fe_add(fe *h, const fe *f, const fe *g) {
    for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
    fe_carry(h);
}
/* h = f - g, computed as f + 8p - g so no limb goes negative (limbs stay < 2^52 on input) */
static inline void fe_sub(fe *h, const fe *f, const fe *g) {
    h->v[0] = f->v[0] + 8 * (FE_MASK51 - 18) - g->v[0];
    for (int i = 1; i < 5; i++) h->v[i] = f->v[i] + 8 * FE_MASK51 - g->v[i];
    fe_carry(h);
}
static inline void fe_neg(fe *h, const fe *f) { fe z; fe_0(&z); fe_sub(h, &z, f); }

static inline void fe_mul(fe *h, const fe *f, const fe *g) {
    const uint64_t *a = f->v, *b = g->v;
    uint64_t b1 = 19 * b[1], b2 = 19 * b[2], b3 = 19 * b[3], b4 = 19 * b[4];
    u128 r0 = (u128)a[0] * b[0] + (u128)a[1] * b4 + (u128)a[2] * b3 + (u128)a[3] * b2 + (u128)a[4] * b1;
    u128 r1 = (u128)a[0] * b[1] + (u128)a[1] * b[0] + (u128)a[2] * b4 + (u128)a[3] * b3 + (u128)a[4] * b2;
    u128 r2 = (u128)a[0] * b[2] + (u128)a[1] * b[1] + (u128)a[2] * b[0] + (u128)a[3] * b4 + (u128)a[4] * b3;
    u128 r3 = (u128)a[0] * b[3] + (u128)a[1] * b[2] + (u128)a[2] * b[1] + (u128)a[3] * b[0] + (u128)a[4] * b4;
    u128 r4 = (u128)a[0] * b[4] + (u128)a[1] * b[3] + (u128)a[2] * b[2] + (u128)a[3] * b[1] + (u128)a[4] * b[0];
    uint64_t c;
    r1 += (uint64_t)(r0 >> 51); h->v[0] = (uint64_t)r0 & FE_MASK51;
    r2 += (uint64_t)(r1 >> 51); h->v[1] = (uint64_t)r1 & FE_MASK51;
    r3 += (uint64_t)(r2 >> 51); h->v[2] = (uint64_t)r2 & FE_MASK51;
    r4 += (uint64_t)(r3 >> 51); h->v[3] = (uint64_t)r3 & FE_MASK51;
    c = (uint64_t)(r4 >> 51);   h->v[4] = (uint64_t)r4 & FE_MASK51;
    h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= FE_MASK51; h->v[1] += c;
}
static inline void fe_sq(fe *h, const fe *f) { fe_mul(h, f, f); }
static inline void fe_sqn(fe *h, const fe *f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }

/* bytes (little endian, bit 255 ignored -- dalek `FieldElement::from_bytes`) */
static inline void fe_frombytes(fe *h, const uint8_t s[32]) {
    uint64_t w[4];
    memcpy(w, s, 32);
    h->v[0] = w[0] & FE_MASK51;
    h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & FE_MASK51;
    h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & FE_MASK51;
    h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & FE_MASK51;
    h->v[4] = (w[3] >> 12) & FE_MASK51;
}
/* canonical encoding */
static inline void fe_tobytes(uint8_t s[32], const fe *f) {
    fe t = *f;
    fe_carry(&t); fe_carry(&t);
    /* t < 2^255 + small; compute t + 19, if that overflows 2^255 then t >= p */
    uint64_t q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    uint64_t c;
    c = t.v[0] >> 51; t.v[0] &= FE_MASK51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= FE_MASK51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= FE_MASK51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= FE_MASK51; t.v[4] += c;
    t.v[4] &= FE_MASK51;
    uint64_t w[4];
    w[0] = t.v[0] | (t.v[1] << 51);
    w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
    w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
    w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
    memcpy(s, w, 32);
}
static inline int fe_iszero(const fe *f) {
    uint8_t s[32]; fe_tobytes(s, f);
    uint8_t r = 0; for (int i = 0; i < 32; i++) r |= s[i];
    return r == 0;
}
static inline int fe_eq(const fe *f, const fe *g) { fe t; fe_sub(&t, f, g); return fe_iszero(&t); }
/* "negative" = low bit of the canonical encoding (RFC 9496 IS_NEGATIVE) */
static inline int fe_isneg(const fe *f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
static inline void fe_cmov(fe *h, const fe *g, int b) { if (b) *h = *g; }
static inline void fe_abs(fe *h, const fe *f) { if (fe_isneg(f)) fe_neg(h, f); else *h = *f; }

/* z^(2^252-3) = z^((p-5)/8) */
static inline void fe_pow22523(fe *out, const fe *z) {
    fe t0, t1, t2;
    fe_sq(&t0, z);                       /* 2 */
    fe_sqn(&t1, &t0, 2);                 /* 8 */
    fe_mul(&t1, z, &t1);                 /* 9 */
    fe_mul(&t0, &t0, &t1);               /* 11 */
    fe_sq(&t0, &t0);                     /* 22 */
    fe_mul(&t0, &t1, &t0);               /* 31 = 2^5-1 */
    fe_sqn(&t1, &t0, 5);  fe_mul(&t0, &t1, &t0);    /* 2^10-1 */
    fe_sqn(&t1, &t0, 10); fe_mul(&t1, &t1, &t0);    /* 2^20-1 */
    fe_sqn(&t2, &t1, 20); fe_mul(&t1, &t2, &t1);    /* 2^40-1 */
    fe_sqn(&t1, &t1, 10); fe_mul(&t0, &t1, &t0);    /* 2^50-1 */
    fe_sqn(&t1, &t0, 50); fe_mul(&t1, &t1, &t0);    /* 2^100-1 */
    fe_sqn(&t2, &t1, 100); fe_mul(&t1, &t2, &t1);   /* 2^200-1 */
    fe_sqn(&t1, &t1, 50); fe_mul(&t0, &t1, &t0);    /* 2^250-1 */
    fe_sqn(&t0, &t0, 2);                 /* 2^252-4 */
    fe_mul(out, &t0, z);                 /* 2^252-3 */
}
/* z^(p-2) = z^(2^255-21): (z^(2^252-3))^8 * z^3 */
static inline void fe_invert(fe *out, const fe *z) {
    fe t, z3;
    fe_pow22523(&t, z);
    fe_sqn(&t, &t, 3);          /* z^(2^255-24) */
    fe_sq(&z3, z); fe_mul(&z3, &z3, z);
    fe_mul(out, &t, &z3);
}
#endif
