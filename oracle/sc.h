/* ORACLE -- TEST INFRASTRUCTURE ONLY (see fe51.h).
 *
 * Integers mod l = 2^252 + 27742317777372353535851937790883648493 (the ristretto255 group
 * order), 32-byte little-endian canonical form = curve25519-dalek-ng `Scalar` (third-party,
 * Cargo.lock:363-364).  Reduction uses 2^252 = -c (mod l) folding on 64-bit limbs -- a
 * different method from the device code (Montgomery on 32-bit limbs).
 */
#ifndef ROFL_ORACLE_SC_H
#define ROFL_ORACLE_SC_H
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t v[4]; } sc;   /* always fully reduced, < l */
typedef unsigned __int128 sc_u128;

static const uint64_t SC_L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0ULL, 0x1000000000000000ULL};
static const uint64_t SC_C[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL};  /* l - 2^252 */

static inline void sc_0(sc *r) { memset(r, 0, sizeof *r); }
static inline void sc_from_u64(sc *r, uint64_t x) { sc_0(r); r->v[0] = x; }
static inline int sc_iszero(const sc *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int sc_eq(const sc *a, const sc *b) { return memcmp(a, b, sizeof *a) == 0; }

/* r = a - b over n limbs, returns borrow */
static inline uint64_t bn_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    uint64_t br = 0;
    for (int i = 0; i < n; i++) {
        sc_u128 t = (sc_u128)a[i] - b[i] - br;
        r[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1;
    }
    return br;
}
static inline uint64_t bn_add(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    uint64_t c = 0;
    for (int i = 0; i < n; i++) {
        sc_u128 t = (sc_u128)a[i] + b[i] + c;
        r[i] = (uint64_t)t; c = (uint64_t)(t >> 64);
    }
    return c;
}
static inline int bn_geq(const uint64_t *a, const uint64_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 1;
}
/* r[na+nb] = a[na] * b[nb] */
static inline void bn_mul(uint64_t *r, const uint64_t *a, int na, const uint64_t *b, int nb) {
    memset(r, 0, 8 * (na + nb));
    for (int i = 0; i < na; i++) {
        uint64_t c = 0;
        for (int j = 0; j < nb; j++) {
            sc_u128 t = (sc_u128)a[i] * b[j] + r[i + j] + c;
            r[i + j] = (uint64_t)t; c = (uint64_t)(t >> 64);
        }
        r[i + nb] = c;
    }
}
/* split x (n limbs) at bit 252: lo[4] = x mod 2^252, hi[n-3] = x >> 252 */
static inline void bn_split252(uint64_t lo[4], uint64_t *hi, const uint64_t *x, int n) {
    for (int i = 0; i < 4; i++) lo[i] = i < n ? x[i] : 0;
    lo[3] &= 0x0fffffffffffffffULL;
    for (int i = 0; i < n - 3; i++) {
        uint64_t a = x[i + 3] >> 60;
        uint64_t b = (i + 4 < n) ? (x[i + 4] << 4) : 0;
        hi[i] = a | b;
    }
}
/* reduce a 512-bit little-endian integer (8 limbs) mod l */
static inline void sc_reduce512(sc *r, const uint64_t x[8]) {
    /* x = lo + 2^252 hi  ==  lo - c*hi ; iterate on c*hi which shrinks by ~127 bits each time */
    uint64_t lo0[4], hi0[5];  bn_split252(lo0, hi0, x, 8);          /* hi0 < 2^260 */
    uint64_t y[7];            bn_mul(y, hi0, 5, SC_C, 2);            /* < 2^385 */
    uint64_t lo1[4], hi1[4];  bn_split252(lo1, hi1, y, 7);          /* hi1 < 2^133 */
    uint64_t z[6];            bn_mul(z, hi1, 4, SC_C, 2);            /* < 2^258 */
    uint64_t lo2[4], hi2[3];  bn_split252(lo2, hi2, z, 6);          /* hi2 < 2^6 */
    uint64_t w[5];            bn_mul(w, hi2, 3, SC_C, 2);            /* < 2^131 */
    /* x == lo0 - lo1 + lo2 - w  (mod l); every term < 2^252.  acc = lo0 + lo2 + 2l - lo1 - w */
    uint64_t acc[5] = {0}, t[5] = {0};
    memcpy(acc, lo0, 32);
    memcpy(t, lo2, 32);   bn_add(acc, acc, t, 5);
    memcpy(t, SC_L, 32);  t[4] = 0; bn_add(acc, acc, t, 5); bn_add(acc, acc, t, 5);
    memcpy(t, lo1, 32);   t[4] = 0; bn_sub(acc, acc, t, 5);
    t[0] = w[0]; t[1] = w[1]; t[2] = w[2]; t[3] = 0; t[4] = 0; bn_sub(acc, acc, t, 5);
    uint64_t l5[5]; memcpy(l5, SC_L, 32); l5[4] = 0;
    while (bn_geq(acc, l5, 5)) bn_sub(acc, acc, l5, 5);
    memcpy(r->v, acc, 32);
}
static inline void sc_from_bytes_wide(sc *r, const uint8_t s[64]) { uint64_t x[8]; memcpy(x, s, 64); sc_reduce512(r, x); }
static inline void sc_from_bytes_mod_order(sc *r, const uint8_t s[32]) {
    uint64_t x[8] = {0}; memcpy(x, s, 32); sc_reduce512(r, x);
}
/* dalek `Scalar::from_canonical_bytes`: 1 if s < l */
static inline int sc_from_canonical_bytes(sc *r, const uint8_t s[32]) {
    uint64_t x[4]; memcpy(x, s, 32);
    if (bn_geq(x, SC_L, 4)) return 0;
    memcpy(r->v, x, 32); return 1;
}
static inline void sc_tobytes(uint8_t s[32], const sc *a) { memcpy(s, a->v, 32); }
static inline void sc_add(sc *r, const sc *a, const sc *b) {
    uint64_t t[4]; bn_add(t, a->v, b->v, 4);       /* < 2^254, no carry out */
    if (bn_geq(t, SC_L, 4)) bn_sub(t, t, SC_L, 4);
    memcpy(r->v, t, 32);
}
static inline void sc_sub(sc *r, const sc *a, const sc *b) {
    uint64_t t[4];
    if (bn_sub(t, a->v, b->v, 4)) bn_add(t, t, SC_L, 4);
    memcpy(r->v, t, 32);
}
static inline void sc_neg(sc *r, const sc *a) { sc z; sc_0(&z); sc_sub(r, &z, a); }
static inline void sc_mul(sc *r, const sc *a, const sc *b) {
    uint64_t x[8]; bn_mul(x, a->v, 4, b->v, 4); sc_reduce512(r, x);
}
/* r = a*b + c */
static inline void sc_muladd(sc *r, const sc *a, const sc *b, const sc *c) { sc t; sc_mul(&t, a, b); sc_add(r, &t, c); }
/* a^(l-2) */
static inline void sc_invert(sc *r, const sc *a) {
    uint64_t e[4]; memcpy(e, SC_L, 32); e[0] -= 2;
    sc acc; sc_from_u64(&acc, 1);
    for (int i = 252; i >= 0; i--) {
        sc_mul(&acc, &acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) sc_mul(&acc, &acc, a);
    }
    *r = acc;
}
static inline void sc_pow_u64(sc *r, const sc *a, uint64_t e) {
    sc acc, base = *a; sc_from_u64(&acc, 1);
    while (e) { if (e & 1) sc_mul(&acc, &acc, &base); sc_mul(&base, &base, &base); e >>= 1; }
    *r = acc;
}
#endif
