/* ORACLE -- TEST INFRASTRUCTURE ONLY (see fe51.h).
 *
 * Twisted Edwards curve -x^2 + y^2 = 1 + d x^2 y^2 in extended coordinates and the
 * ristretto255 quotient group (RFC 9496), i.e. curve25519-dalek-ng 4.1.1 `EdwardsPoint` /
 * `RistrettoPoint` / `CompressedRistretto` (third-party).
 * Reference call sites: pedersen_ops.rs:9-25 (commit), el_gamal.rs:31-69,
 * range_proof_vec/mod.rs:237-246 (compress / decompress).
 */
#ifndef ROFL_ORACLE_GE_H
#define ROFL_ORACLE_GE_H
#include "fe51.h"
#include "sc.h"
#include "constants51.h"

typedef struct { fe X, Y, Z, T; } ge;

static inline const fe *K(const uint64_t k[5]) { return (const fe *)k; }

static inline void ge_identity(ge *p) { fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T); }
static inline void ge_base(ge *p) { p->X = *K(K_BASE_X); p->Y = *K(K_BASE_Y); fe_1(&p->Z); p->T = *K(K_BASE_T); }
static inline void ge_neg(ge *r, const ge *p) { fe_neg(&r->X, &p->X); r->Y = p->Y; r->Z = p->Z; fe_neg(&r->T, &p->T); }

/* complete unified addition (add-2008-hwcd-3, a = -1): 8M + 1 mul by 2d */
static inline void ge_add(ge *r, const ge *p, const ge *q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(&a, &p->Y, &p->X); fe_sub(&t, &q->Y, &q->X); fe_mul(&a, &a, &t);
    fe_add(&b, &p->Y, &p->X); fe_add(&t, &q->Y, &q->X); fe_mul(&b, &b, &t);
    fe_mul(&c, &p->T, &q->T); fe_mul(&c, &c, K(K_D2));
    fe_mul(&d, &p->Z, &q->Z); fe_add(&d, &d, &d);
    fe_sub(&e, &b, &a); fe_sub(&f, &d, &c); fe_add(&g, &d, &c); fe_add(&h, &b, &a);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->T, &e, &h); fe_mul(&r->Z, &f, &g);
}
static inline void ge_sub(ge *r, const ge *p, const ge *q) { ge n; ge_neg(&n, q); ge_add(r, p, &n); }
/* dbl-2008-hwcd with a = -1 */
static inline void ge_dbl(ge *r, const ge *p) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sq(&a, &p->X); fe_sq(&b, &p->Y); fe_sq(&c, &p->Z); fe_add(&c, &c, &c);
    fe_neg(&d, &a);                                   /* D = a*A = -A */
    fe_add(&t, &p->X, &p->Y); fe_sq(&t, &t); fe_sub(&e, &t, &a); fe_sub(&e, &e, &b);
    fe_add(&g, &d, &b); fe_sub(&f, &g, &c); fe_sub(&h, &d, &b);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->T, &e, &h); fe_mul(&r->Z, &f, &g);
}
/* ristretto equality: X1 Y2 == Y1 X2  or  Y1 Y2 == X1 X2 */
static inline int ge_eq(const ge *p, const ge *q) {
    fe a, b;
    fe_mul(&a, &p->X, &q->Y); fe_mul(&b, &p->Y, &q->X); if (fe_eq(&a, &b)) return 1;
    fe_mul(&a, &p->Y, &q->Y); fe_mul(&b, &p->X, &q->X); return fe_eq(&a, &b);
}
static inline int ge_is_identity(const ge *p) { ge id; ge_identity(&id); return ge_eq(p, &id); }

/* RFC 9496 4.2 SQRT_RATIO_M1: returns was_square, r = sqrt(u/v) or sqrt(i*u/v), r non-negative */
static inline int fe_sqrt_ratio_m1(fe *r, const fe *u, const fe *v) {
    fe v3, v7, t, check, neg_u, neg_u_i;
    fe_sq(&v3, v); fe_mul(&v3, &v3, v);            /* v^3 */
    fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);          /* v^7 */
    fe_mul(&t, u, &v7); fe_pow22523(&t, &t);       /* (u v^7)^((p-5)/8) */
    fe_mul(&t, &t, &v3); fe_mul(&t, &t, u);        /* r = u v^3 (u v^7)^((p-5)/8) */
    fe_sq(&check, &t); fe_mul(&check, &check, v);  /* v r^2 */
    fe_neg(&neg_u, u); fe_mul(&neg_u_i, &neg_u, K(K_SQRT_M1));
    int correct = fe_eq(&check, u), flipped = fe_eq(&check, &neg_u), flipped_i = fe_eq(&check, &neg_u_i);
    if (flipped || flipped_i) fe_mul(&t, &t, K(K_SQRT_M1));
    fe_abs(r, &t);
    return correct || flipped;
}

/* RFC 9496 4.3.2 Encode */
static inline void ge_compress(uint8_t s[32], const ge *p) {
    fe u1, u2, u2sq, inv, den1, den2, zinv, ix, iy, ench, x, y, den_inv, t, one;
    fe_1(&one);
    fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t);
    fe_mul(&u2, &p->X, &p->Y);
    fe_sq(&u2sq, &u2); fe_mul(&t, &u1, &u2sq);
    fe_sqrt_ratio_m1(&inv, &one, &t);
    fe_mul(&den1, &inv, &u1); fe_mul(&den2, &inv, &u2);
    fe_mul(&zinv, &den1, &den2); fe_mul(&zinv, &zinv, &p->T);
    fe_mul(&ix, &p->X, K(K_SQRT_M1)); fe_mul(&iy, &p->Y, K(K_SQRT_M1));
    fe_mul(&ench, &den1, K(K_INVSQRT_A_MINUS_D));
    fe_mul(&t, &p->T, &zinv);
    int rotate = fe_isneg(&t);
    x = p->X; y = p->Y; den_inv = den2;
    if (rotate) { x = iy; y = ix; den_inv = ench; }
    fe_mul(&t, &x, &zinv);
    if (fe_isneg(&t)) fe_neg(&y, &y);
    fe_sub(&t, &p->Z, &y); fe_mul(&t, &t, &den_inv);
    fe_abs(&t, &t);
    fe_tobytes(s, &t);
}
/* RFC 9496 4.3.1 Decode; returns 1 on success */
static inline int ge_decompress(ge *p, const uint8_t s[32]) {
    fe sfe, ss, u1, u2, u2sq, v, inv, denx, deny, t, one;
    uint8_t chk[32];
    fe_frombytes(&sfe, s); fe_tobytes(chk, &sfe);
    if (memcmp(chk, s, 32) != 0) return 0;          /* non-canonical (also rejects bit 255 set) */
    if (s[0] & 1) return 0;                         /* negative */
    fe_1(&one);
    fe_sq(&ss, &sfe);
    fe_sub(&u1, &one, &ss); fe_add(&u2, &one, &ss); fe_sq(&u2sq, &u2);
    fe_mul(&v, K(K_D), &u1); fe_mul(&v, &v, &u1); fe_neg(&v, &v); fe_sub(&v, &v, &u2sq);   /* -(d u1^2) - u2^2 */
    fe_mul(&t, &v, &u2sq);
    int ok = fe_sqrt_ratio_m1(&inv, &one, &t);
    fe_mul(&denx, &inv, &u2); fe_mul(&deny, &inv, &denx); fe_mul(&deny, &deny, &v);
    fe_add(&t, &sfe, &sfe); fe_mul(&t, &t, &denx); fe_abs(&p->X, &t);
    fe_mul(&p->Y, &u1, &deny);
    fe_1(&p->Z);
    fe_mul(&p->T, &p->X, &p->Y);
    if (!ok || fe_isneg(&p->T) || fe_iszero(&p->Y)) return 0;
    return 1;
}
/* RFC 9496 4.3.4 MAP (dalek `elligator_ristretto_flavor`) */
static inline void ge_elligator(ge *p, const fe *r0) {
    fe r, ns, c, dd, s, sp, nt, w0, w1, w2, w3, t, one;
    fe_1(&one);
    fe_sq(&r, r0); fe_mul(&r, &r, K(K_SQRT_M1));
    fe_add(&ns, &r, &one); fe_mul(&ns, &ns, K(K_ONE_MINUS_D_SQ));         /* u */
    fe_neg(&c, &one);
    fe_mul(&t, &r, K(K_D)); fe_sub(&dd, &c, &t);                          /* c - d r */
    fe_add(&t, &r, K(K_D)); fe_mul(&dd, &dd, &t);                         /* v = (c - d r)(r + d) */
    int was_square = fe_sqrt_ratio_m1(&s, &ns, &dd);
    fe_mul(&sp, &s, r0); fe_abs(&sp, &sp); fe_neg(&sp, &sp);              /* s' = -|s r0| */
    if (!was_square) { s = sp; c = r; }
    fe_sub(&t, &r, &one); fe_mul(&nt, &c, &t); fe_mul(&nt, &nt, K(K_D_MINUS_ONE_SQ)); fe_sub(&nt, &nt, &dd);
    fe_add(&w0, &s, &s); fe_mul(&w0, &w0, &dd);
    fe_mul(&w1, &nt, K(K_SQRT_AD_MINUS_ONE));
    fe_sq(&t, &s); fe_sub(&w2, &one, &t); fe_add(&w3, &one, &t);
    fe_mul(&p->X, &w0, &w3); fe_mul(&p->Y, &w2, &w1); fe_mul(&p->Z, &w1, &w3); fe_mul(&p->T, &w0, &w2);
}
/* dalek `RistrettoPoint::from_uniform_bytes` == libsodium crypto_core_ristretto255_from_hash */
static inline void ge_from_uniform_bytes(ge *p, const uint8_t b[64]) {
    fe r1, r2; ge p1, p2;
    fe_frombytes(&r1, b); fe_frombytes(&r2, b + 32);
    ge_elligator(&p1, &r1); ge_elligator(&p2, &r2);
    ge_add(p, &p1, &p2);
}

/* ---- scalar multiplication --------------------------------------------------------------- */
/* signed radix-16 digits of a 253-bit scalar (64 digits in [-8,8]) */
static inline void sc_radix16(int8_t e[64], const sc *a) {
    uint8_t s[32]; sc_tobytes(s, a);
    for (int i = 0; i < 32; i++) { e[2 * i] = s[i] & 15; e[2 * i + 1] = (s[i] >> 4) & 15; }
    int8_t carry = 0;
    for (int i = 0; i < 63; i++) { e[i] += carry; carry = (e[i] + 8) >> 4; e[i] -= carry << 4; }
    e[63] += carry;
}
/* width-w non-adjacent form, 256 digits */
static inline void sc_naf(int8_t naf[256], const sc *a, int w) {
    uint64_t x[5] = {a->v[0], a->v[1], a->v[2], a->v[3], 0};
    memset(naf, 0, 256);
    int width = 1 << w, mask = width - 1, pos = 0, carry = 0;
    while (pos < 256) {
        int idx = pos >> 6, bit = pos & 63;
        uint64_t buf = bit < 64 - w ? x[idx] >> bit : (x[idx] >> bit) | (x[idx + 1] << (64 - bit));
        int window = carry + (int)(buf & mask);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)(window - width); }
        pos += w;
    }
}
/* variable-base scalar multiplication, signed radix-16 with a table of 1..8 multiples
 * (the shape of dalek's constant-time `EdwardsPoint * Scalar`) */
static inline void ge_scalarmult(ge *r, const sc *a, const ge *p) {
    ge tab[9]; int8_t e[64];
    ge_identity(&tab[0]); tab[1] = *p;
    for (int i = 2; i <= 8; i++) ge_add(&tab[i], &tab[i - 1], p);
    sc_radix16(e, a);
    ge acc; ge_identity(&acc);
    for (int i = 63; i >= 0; i--) {
        if (i != 63) { ge_dbl(&acc, &acc); ge_dbl(&acc, &acc); ge_dbl(&acc, &acc); ge_dbl(&acc, &acc); }
        if (e[i] > 0) ge_add(&acc, &acc, &tab[e[i]]);
        else if (e[i] < 0) ge_sub(&acc, &acc, &tab[-e[i]]);
    }
    *r = acc;
}
#endif
