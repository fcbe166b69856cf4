// Twisted Edwards -x^2+y^2 = 1+d x^2 y^2 (extended coordinates, a = -1) and ristretto255 encode/decode/
// Elligator (RFC 9496) on the radix-2^25.5 field.  curve25519-dalek-ng `EdwardsPoint`/`RistrettoPoint`
// semantics; reference call sites pedersen_ops.rs:9-25, el_gamal.rs:31-69, range_proof_vec/mod.rs:237-246.
// Formulas: Hisil-Wong-Carter-Dawson 2008 (add-2008-hwcd-3 / dbl-2008-hwcd).  All __host__ __device__.
//
// Representations:
//   ge_p3     (X:Y:Z:T), x=X/Z y=Y/Z xy=T/Z     -- the working form; every coordinate REDUCED (scale 1)
//   ge_p1p1   completed point ((X:Z),(Y:T))      -- output of add/dbl before the final 3-4 multiplications
//   ge_niels  affine (y+x, y-x, 2dxy)            -- fixed generators (tables in HBM), 7M mixed add
//   ge_cached (Y+X, Y-X, Z, 2dT)                 -- variable points in window tables, 8M add
#pragma once
#include "fe25519.cuh"
#include "sc25519.cuh"
#include "constants32.cuh"

struct ge_p3 { fe X, Y, Z, T; };
struct ge_p2 { fe X, Y, Z; };
struct ge_p1p1 { fe X, Y, Z, T; };
struct ge_niels { fe yplusx, yminusx, xy2d; };
struct ge_cached { fe YplusX, YminusX, Z, T2d; };

#if defined(__CUDA_ARCH__)
#define GE_CONST static __device__ __constant__ const
#else
#define GE_CONST static const
#endif
GE_CONST fe GE_D = {FE_D};
GE_CONST fe GE_D2 = {FE_D2};
GE_CONST fe GE_SQRT_M1 = {FE_SQRT_M1};
GE_CONST fe GE_SQRT_AD_MINUS_ONE = {FE_SQRT_AD_MINUS_ONE};
GE_CONST fe GE_INVSQRT_A_MINUS_D = {FE_INVSQRT_A_MINUS_D};
GE_CONST fe GE_ONE_MINUS_D_SQ = {FE_ONE_MINUS_D_SQ};
GE_CONST fe GE_D_MINUS_ONE_SQ = {FE_D_MINUS_ONE_SQ};
GE_CONST fe GE_BASE_X = {FE_BASE_X};
GE_CONST fe GE_BASE_Y = {FE_BASE_Y};
GE_CONST fe GE_BASE_T = {FE_BASE_T};

HD void ge_p3_0(ge_p3 &p) { fe_0(p.X); fe_1(p.Y); fe_1(p.Z); fe_0(p.T); }
HD void ge_base(ge_p3 &p) { p.X = GE_BASE_X; p.Y = GE_BASE_Y; fe_1(p.Z); p.T = GE_BASE_T; }
HD void ge_niels_0(ge_niels &n) { fe_1(n.yplusx); fe_1(n.yminusx); fe_0(n.xy2d); }

// completed -> projective/extended.  Operand order matters for the limb-scale rules of fe_mul (second operand
// scale <= 3): producers must deliver scale(X) <= 3, scale(Y) <= 3, scale(Z) <= 3; T may be up to 5.
HD void ge_p1p1_to_p2(ge_p2 &r, const ge_p1p1 &p) { fe_mul(r.X, p.T, p.X); fe_mul(r.Y, p.Y, p.Z); fe_mul(r.Z, p.T, p.Z); }
HD void ge_p1p1_to_p3(ge_p3 &r, const ge_p1p1 &p) { fe_mul(r.X, p.T, p.X); fe_mul(r.Y, p.Y, p.Z); fe_mul(r.Z, p.T, p.Z); fe_mul(r.T, p.X, p.Y); }

// doubling: 4S; inputs reduced.  Raw output scales: X 5, Y 2, Z 3, T 5; ge_dbl_fix carries X down to 1.
HD void ge_dbl_p1p1(ge_p1p1 &r, const fe &X, const fe &Y, const fe &Z) {
    fe xx, yy, zz, xy;
    fe_sq(xx, X); fe_sq(yy, Y); fe_sq(zz, Z);
    fe_add(xy, X, Y); fe_sq(xy, xy);                 // (X+Y)^2, input scale 2
    fe_add(r.Y, yy, xx);                             // Y3 = YY+XX          scale 2
    fe_sub(r.Z, yy, xx);                             // Z3 = YY-XX          scale 3
    fe_sub4(r.X, xy, r.Y);                           // X3 = (X+Y)^2 - Y3   scale 5
    fe t; fe_add(t, zz, zz); fe_add(t, t, xx);       // 2ZZ + XX            scale 3
    fe_sub(r.T, t, yy);                              // T3 = 2ZZ - Z3       scale 5
}
HD void ge_dbl_fix(ge_p1p1 &r) { fe_carry(r.X, r.X); }

HD void ge_p3_dbl(ge_p3 &r, const ge_p3 &p) { ge_p1p1 t; ge_dbl_p1p1(t, p.X, p.Y, p.Z); ge_dbl_fix(t); ge_p1p1_to_p3(r, t); }
HD void ge_p2_dbl(ge_p2 &r, const ge_p2 &p) { ge_p1p1 t; ge_dbl_p1p1(t, p.X, p.Y, p.Z); ge_dbl_fix(t); ge_p1p1_to_p2(r, t); }

// mixed addition p3 + niels (7M): A=(Y1-X1)(y2-x2) B=(Y1+X1)(y2+x2) C=T1*2d*x2y2 D=2Z1
HD void ge_madd_p1p1(ge_p1p1 &r, const ge_p3 &p, const ge_niels &q) {
    fe a, b, c, d;
    fe_add(b, p.Y, p.X); fe_mul(b, b, q.yplusx);
    fe_sub(a, p.Y, p.X); fe_mul(a, a, q.yminusx);
    fe_mul(c, p.T, q.xy2d);
    fe_add(d, p.Z, p.Z);                             // scale 2
    fe_sub(r.X, b, a);                               // E = B-A   scale 3
    fe_add(r.Y, b, a);                               // H = B+A   scale 2
    fe_add(r.Z, d, c);                               // G = D+C   scale 3
    fe_sub(r.T, d, c);                               // F = D-C   scale 4 (first operand only)
}
HD void ge_msub_p1p1(ge_p1p1 &r, const ge_p3 &p, const ge_niels &q) {
    fe a, b, c, d;
    fe_add(b, p.Y, p.X); fe_mul(b, b, q.yminusx);
    fe_sub(a, p.Y, p.X); fe_mul(a, a, q.yplusx);
    fe_mul(c, p.T, q.xy2d);
    fe_add(d, p.Z, p.Z);
    fe_sub(r.X, b, a); fe_add(r.Y, b, a);
    fe_sub(r.Z, d, c); fe_carry(r.Z, r.Z);           // G = D-C
    fe_add(r.T, d, c);                               // F = D+C   scale 3
}
HD void ge_madd(ge_p3 &r, const ge_p3 &p, const ge_niels &q) { ge_p1p1 t; ge_madd_p1p1(t, p, q); ge_p1p1_to_p3(r, t); }
HD void ge_msub(ge_p3 &r, const ge_p3 &p, const ge_niels &q) { ge_p1p1 t; ge_msub_p1p1(t, p, q); ge_p1p1_to_p3(r, t); }
// add / subtract by sign flag with a branch-free operand swap: lanes of a warp with different signs run ONE addition
HD void ge_madd_signed(ge_p3 &r, const ge_p3 &p, const ge_niels &q, bool neg) {
    ge_niels n; n.yplusx = q.yplusx; n.yminusx = q.yminusx; n.xy2d = q.xy2d;
    fe nx; fe_neg(nx, q.xy2d); fe_carry(nx, nx);
    fe_cmov(n.yplusx, q.yminusx, neg); fe_cmov(n.yminusx, q.yplusx, neg); fe_cmov(n.xy2d, nx, neg);
    ge_madd(r, p, n);
}

HD void ge_p3_to_cached(ge_cached &r, const ge_p3 &p) {
    fe_add(r.YplusX, p.Y, p.X); fe_carry(r.YplusX, r.YplusX);
    fe_sub(r.YminusX, p.Y, p.X); fe_carry(r.YminusX, r.YminusX);
    r.Z = p.Z; fe_mul(r.T2d, p.T, GE_D2);
}
// p3 + cached (8M)
HD void ge_add_p1p1(ge_p1p1 &r, const ge_p3 &p, const ge_cached &q) {
    fe a, b, c, d;
    fe_add(b, p.Y, p.X); fe_mul(b, b, q.YplusX);
    fe_sub(a, p.Y, p.X); fe_mul(a, a, q.YminusX);
    fe_mul(c, p.T, q.T2d);
    fe_mul(d, p.Z, q.Z); fe_add(d, d, d);
    fe_sub(r.X, b, a); fe_add(r.Y, b, a);
    fe_add(r.Z, d, c);
    fe_sub(r.T, d, c);
}
HD void ge_sub_p1p1(ge_p1p1 &r, const ge_p3 &p, const ge_cached &q) {
    fe a, b, c, d;
    fe_add(b, p.Y, p.X); fe_mul(b, b, q.YminusX);
    fe_sub(a, p.Y, p.X); fe_mul(a, a, q.YplusX);
    fe_mul(c, p.T, q.T2d);
    fe_mul(d, p.Z, q.Z); fe_add(d, d, d);
    fe_sub(r.X, b, a); fe_add(r.Y, b, a);
    fe_sub(r.Z, d, c); fe_carry(r.Z, r.Z);
    fe_add(r.T, d, c);
}
HD void ge_add_cached(ge_p3 &r, const ge_p3 &p, const ge_cached &q) { ge_p1p1 t; ge_add_p1p1(t, p, q); ge_p1p1_to_p3(r, t); }
HD void ge_sub_cached(ge_p3 &r, const ge_p3 &p, const ge_cached &q) { ge_p1p1 t; ge_sub_p1p1(t, p, q); ge_p1p1_to_p3(r, t); }
HD void ge_add(ge_p3 &r, const ge_p3 &p, const ge_p3 &q) { ge_cached c; ge_p3_to_cached(c, q); ge_add_cached(r, p, c); }
HD void ge_sub(ge_p3 &r, const ge_p3 &p, const ge_p3 &q) { ge_cached c; ge_p3_to_cached(c, q); ge_sub_cached(r, p, c); }
HD void ge_add_cached_signed(ge_p3 &r, const ge_p3 &p, const ge_cached &q, bool neg) {
    ge_cached n; n.YplusX = q.YplusX; n.YminusX = q.YminusX; n.Z = q.Z; n.T2d = q.T2d;
    fe nt; fe_neg(nt, q.T2d); fe_carry(nt, nt);
    fe_cmov(n.YplusX, q.YminusX, neg); fe_cmov(n.YminusX, q.YplusX, neg); fe_cmov(n.T2d, nt, neg);
    ge_add_cached(r, p, n);
}
HD void ge_neg(ge_p3 &r, const ge_p3 &p) { fe_neg(r.X, p.X); fe_carry(r.X, r.X); r.Y = p.Y; r.Z = p.Z; fe_neg(r.T, p.T); fe_carry(r.T, r.T); }

// ristretto equality (X1 Y2 == Y1 X2  or  Y1 Y2 == X1 X2)
HD bool ge_eq(const ge_p3 &p, const ge_p3 &q) {
    fe a, b; fe_mul(a, p.X, q.Y); fe_mul(b, p.Y, q.X); bool e1 = fe_eq(a, b);
    fe_mul(a, p.Y, q.Y); fe_mul(b, p.X, q.X); bool e2 = fe_eq(a, b);
    return e1 || e2;
}
// identity coset: X == 0 or Y == 0 ... as ristretto: equal to (0,1)
HD bool ge_is_identity(const ge_p3 &p) { return fe_iszero(p.X) || fe_iszero(p.Y); }

// RFC 9496 SQRT_RATIO_M1; u, v scale <= 2; r reduced, non-negative
HDNI bool fe_sqrt_ratio_m1(fe &r, const fe &u, const fe &v) {
    fe v3, v7, t, check, uc, neg_u, neg_u_i;
    fe_sq(v3, v); fe_mul(v3, v3, v);
    fe_sq(v7, v3); fe_mul(v7, v7, v);
    fe_mul(t, u, v7); fe_pow22523(t, t);
    fe_mul(t, t, v3); fe_mul(t, t, u);
    fe_sq(check, t); fe_mul(check, check, v);
    fe_carry(uc, u); fe_neg(neg_u, uc); fe_carry(neg_u, neg_u); fe_mul(neg_u_i, neg_u, GE_SQRT_M1);
    bool correct = fe_eq(check, uc), flipped = fe_eq(check, neg_u), flipped_i = fe_eq(check, neg_u_i);
    fe ti; fe_mul(ti, t, GE_SQRT_M1);
    fe_cmov(t, ti, flipped || flipped_i);
    fe_abs(r, t);
    return correct || flipped;
}

// RFC 9496 4.3.2 Encode
HDNI void ge_compress(uint8_t s[32], const ge_p3 &p) {
    fe u1, u2, u2sq, inv, den1, den2, zinv, ix, iy, ench, x, y, den_inv, t, one;
    fe_1(one);
    fe_add(u1, p.Z, p.Y); fe_sub(t, p.Z, p.Y); fe_mul(u1, t, u1);
    fe_mul(u2, p.X, p.Y);
    fe_sq(u2sq, u2); fe_mul(t, u1, u2sq);
    fe_sqrt_ratio_m1(inv, one, t);
    fe_mul(den1, inv, u1); fe_mul(den2, inv, u2);
    fe_mul(zinv, den1, den2); fe_mul(zinv, zinv, p.T);
    fe_mul(ix, p.X, GE_SQRT_M1); fe_mul(iy, p.Y, GE_SQRT_M1);
    fe_mul(ench, den1, GE_INVSQRT_A_MINUS_D);
    fe_mul(t, p.T, zinv);
    bool rotate = fe_isneg(t);
    x = p.X; y = p.Y; den_inv = den2;
    fe_cmov(x, iy, rotate); fe_cmov(y, ix, rotate); fe_cmov(den_inv, ench, rotate);
    fe_mul(t, x, zinv);
    fe ny; fe_neg(ny, y); fe_carry(ny, ny);
    fe_cmov(y, ny, fe_isneg(t));
    fe_sub(t, p.Z, y); fe_mul(t, t, den_inv);
    fe_abs(t, t);
    fe_tobytes(s, t);
}
// RFC 9496 4.3.1 Decode
HDNI bool ge_decompress(ge_p3 &p, const uint8_t s[32]) {
    fe sfe, ss, u1, u2, u2sq, v, inv, denx, deny, t, one;
    uint8_t chk[32];
    fe_frombytes(sfe, s); fe_tobytes(chk, sfe);
    bool canon = true; for (int i = 0; i < 32; i++) canon = canon && (chk[i] == s[i]);
    bool neg = s[0] & 1;
    fe_1(one);
    fe_sq(ss, sfe);
    fe_sub(u1, one, ss); fe_carry(u1, u1); fe_add(u2, one, ss); fe_sq(u2sq, u2);
    fe_sq(v, u1); fe_mul(v, v, GE_D); fe_neg(v, v); fe_carry(v, v); fe_sub(v, v, u2sq); fe_carry(v, v);
    fe_mul(t, v, u2sq);
    bool ok = fe_sqrt_ratio_m1(inv, one, t);
    fe_mul(denx, inv, u2); fe_mul(deny, inv, denx); fe_mul(deny, deny, v);
    fe_add(t, sfe, sfe); fe_mul(t, t, denx); fe_abs(p.X, t);
    fe_mul(p.Y, u1, deny);
    fe_1(p.Z);
    fe_mul(p.T, p.X, p.Y);
    return canon && !neg && ok && !fe_isneg(p.T) && !fe_iszero(p.Y);
}
// RFC 9496 4.3.4 MAP
HDNI void ge_elligator(ge_p3 &p, const fe &r0) {
    fe r, ns, c, dd, s, sp, nt, w0, w1, w2, w3, t, one, mone;
    fe_1(one); fe_neg(mone, one); fe_carry(mone, mone);
    fe_sq(r, r0); fe_mul(r, r, GE_SQRT_M1);
    fe_add(ns, r, one); fe_mul(ns, ns, GE_ONE_MINUS_D_SQ);
    fe_mul(t, r, GE_D); fe_sub(dd, mone, t); fe_carry(dd, dd);
    fe_add(t, r, GE_D); fe_mul(dd, dd, t);
    bool was_square = fe_sqrt_ratio_m1(s, ns, dd);
    fe_mul(sp, s, r0); fe_abs(sp, sp); fe_neg(sp, sp); fe_carry(sp, sp);
    c = mone;
    fe_cmov(s, sp, !was_square); fe_cmov(c, r, !was_square);
    fe_sub(t, r, one); fe_carry(t, t); fe_mul(nt, c, t); fe_mul(nt, nt, GE_D_MINUS_ONE_SQ); fe_sub(nt, nt, dd); fe_carry(nt, nt);
    fe_add(w0, s, s); fe_mul(w0, w0, dd);
    fe_mul(w1, nt, GE_SQRT_AD_MINUS_ONE);
    fe_sq(t, s); fe_sub(w2, one, t); fe_carry(w2, w2); fe_add(w3, one, t);
    fe_mul(p.X, w0, w3); fe_mul(p.Y, w2, w1); fe_mul(p.Z, w1, w3); fe_mul(p.T, w0, w2);
}
HD void ge_from_uniform_bytes(ge_p3 &p, const uint8_t b[64]) {
    fe r1, r2; ge_p3 p1, p2;
    fe_frombytes(r1, b); fe_frombytes(r2, b + 32);
    ge_elligator(p1, r1); ge_elligator(p2, r2);
    ge_add(p, p1, p2);
}
// affine niels form of a p3 point given 1/Z
HD void ge_p3_to_niels(ge_niels &n, const ge_p3 &p, const fe &zinv) {
    fe x, y, xy;
    fe_mul(x, p.X, zinv); fe_mul(y, p.Y, zinv);
    fe_add(n.yplusx, y, x); fe_carry(n.yplusx, n.yplusx);
    fe_sub(n.yminusx, y, x); fe_carry(n.yminusx, n.yminusx);
    fe_mul(xy, x, y); fe_mul(n.xy2d, xy, GE_D2);
}
HD void ge_niels_to_p3(ge_p3 &p, const ge_niels &n) {
    // y = (yplusx + yminusx)/2, x = (yplusx - yminusx)/2 : use projective Z = 2
    fe_add(p.Y, n.yplusx, n.yminusx); fe_carry(p.Y, p.Y);
    fe_sub(p.X, n.yplusx, n.yminusx); fe_carry(p.X, p.X);
    fe_0(p.Z); p.Z.v[0] = 2;
    // T = X*Y/Z = 2xy*2... X=2x, Y=2y, Z=2 -> T = XY/Z = 2xy
    fe t; fe_mul(t, p.X, p.Y);        // 4xy
    // divide by 2: multiply by (p+1)/2 is costly; instead scale everything: use Z=4 with X=4x? keep exact:
    // choose X' = 2X = 4x, Y' = 2Y = 4y, Z' = 4, T' = X'Y'/Z' = 4xy = t
    fe_add(p.X, p.X, p.X); fe_add(p.Y, p.Y, p.Y); fe_carry(p.X, p.X); fe_carry(p.Y, p.Y);
    p.Z.v[0] = 4; p.T = t;
}

// ---- scalar digit recoding -----------------------------------------------------------------------
// signed radix-256 digits of a canonical scalar: 32 digits in [-128, 128], sum d_i 256^i
HD void sc_radix256(int16_t d[32], const sc &a) {
    int carry = 0;
    for (int i = 0; i < 32; i++) {
        int b = (int)((a.v[i >> 2] >> (8 * (i & 3))) & 0xff) + carry;
        carry = 0;
        if (b > 128) { b -= 256; carry = 1; }
        d[i] = (int16_t)b;
    }
    // a < 2^253: top byte <= 0x1f, so the final carry is always 0
}
// width-w NAF (w in 2..8): 256 digits, odd, |d| < 2^(w-1)
HDNI void sc_naf(int8_t naf[256], const sc &a, int w) {
    uint32_t x[9]; for (int i = 0; i < 8; i++) x[i] = a.v[i]; x[8] = 0;
    for (int i = 0; i < 256; i++) naf[i] = 0;
    const int width = 1 << w, mask = width - 1;
    int pos = 0, carry = 0;
    while (pos < 256) {
        int idx = pos >> 5, bit = pos & 31;
        uint64_t two = (uint64_t)x[idx] | ((uint64_t)x[idx + 1] << 32);
        int window = carry + (int)((two >> bit) & mask);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < width / 2) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)(window - width); }
        pos += w;
    }
}
