// Per-thread building blocks of the kernels (all __host__ __device__ so tests/hostsim can run them on the CPU):
// fixed-base tables and multiplication, uniform-NAF and per-lane radix-16 variable-base ladders, the generator
// fold step, the BulletproofGens hash chain, the f32 -> fixed-point conversion and the per-element square proof.
#pragma once
#include "ge25519.cuh"
#include "hash.cuh"
#include <math.h>

// ---- HBM record formats (16-byte aligned, one record per point / scalar) ------------------------------------------
struct __align__(16) niels_st { uint32_t w[24]; };     //  96 B: y+x | y-x | 2dxy
struct __align__(16) p3_st { uint32_t w[32]; };        // 128 B: X | Y | Z | T
struct __align__(16) sc_st { uint32_t w[8]; };

HD void ld_fe2(fe &a, fe &b, const uint4 *q) {          // two field elements = four 16-byte loads
    uint4 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3];
    a.v[0] = v0.x; a.v[1] = v0.y; a.v[2] = v0.z; a.v[3] = v0.w; a.v[4] = v1.x; a.v[5] = v1.y; a.v[6] = v1.z; a.v[7] = v1.w;
    b.v[0] = v2.x; b.v[1] = v2.y; b.v[2] = v2.z; b.v[3] = v2.w; b.v[4] = v3.x; b.v[5] = v3.y; b.v[6] = v3.z; b.v[7] = v3.w;
}
HD void ld_fe1(fe &a, const uint4 *q) {
    uint4 v0 = q[0], v1 = q[1];
    a.v[0] = v0.x; a.v[1] = v0.y; a.v[2] = v0.z; a.v[3] = v0.w; a.v[4] = v1.x; a.v[5] = v1.y; a.v[6] = v1.z; a.v[7] = v1.w;
}
HD void st_fe1(uint4 *q, const fe &a) { q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]); q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]); }
HD void ld_niels(ge_niels &n, const niels_st *p) { const uint4 *q = (const uint4 *)p; ld_fe2(n.yplusx, n.yminusx, q); ld_fe1(n.xy2d, q + 4); }
HD void st_niels(niels_st *p, const ge_niels &n) { uint4 *q = (uint4 *)p; st_fe1(q, n.yplusx); st_fe1(q + 2, n.yminusx); st_fe1(q + 4, n.xy2d); }
HD void ld_p3(ge_p3 &r, const p3_st *p) { const uint4 *q = (const uint4 *)p; ld_fe2(r.X, r.Y, q); ld_fe2(r.Z, r.T, q + 4); }
HD void st_p3(p3_st *p, const ge_p3 &r) { uint4 *q = (uint4 *)p; st_fe1(q, r.X); st_fe1(q + 2, r.Y); st_fe1(q + 4, r.Z); st_fe1(q + 6, r.T); }
HD void ld_sc(sc &s, const sc_st *p) { const uint4 *q = (const uint4 *)p; uint4 a = q[0], b = q[1]; s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w; s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w; }
HD void st_sc(sc_st *p, const sc &s) { uint4 *q = (uint4 *)p; q[0] = make_uint4(s.v[0], s.v[1], s.v[2], s.v[3]); q[1] = make_uint4(s.v[4], s.v[5], s.v[6], s.v[7]); }
HD void ld_bytes32(uint8_t b[32], const uint8_t *p) { const uint4 *q = (const uint4 *)p; uint4 a = q[0], c = q[1]; uint32_t w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w}; for (int i = 0; i < 32; i++) b[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3))); }
HD void st_bytes32(uint8_t *p, const uint8_t b[32]) { uint32_t w[8]; for (int i = 0; i < 8; i++) w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24); uint4 *q = (uint4 *)p; q[0] = make_uint4(w[0], w[1], w[2], w[3]); q[1] = make_uint4(w[4], w[5], w[6], w[7]); }



// ---- f32 -> fixed point: Fix::saturating_from_float(|x|).to_bits() (conversion32.rs:11-19, fp.rs:35-137) -------
// round to nearest, ties to even; saturate at 2^n_bits - 1.  returns -1 for NaN (the reference panics).
HD uint64_t fix_max(int n_bits) { return n_bits == 64 ? ~0ULL : ((1ULL << n_bits) - 1); }
HD int f32_to_fix(uint64_t *raw, float x, int n_bits, int frac) {
    if (x != x) return -1;
    float r = rintf(fabsf(x) * (float)(1u << frac));          // exact: power-of-two scaling then RNE
    float lim = (n_bits == 64) ? 18446744073709551616.0f : (float)(1ULL << n_bits);
    *raw = (r >= lim) ? fix_max(n_bits) : (uint64_t)r;
    return 0;
}
HD int f32_to_scalar(sc &s, float x, int n_bits, int frac) {
    uint64_t raw; if (f32_to_fix(&raw, x, n_bits, frac)) return -1;
    sc_from_u64(s, raw); if (x < 0.0f) sc_neg(s, s);
    return 0;
}
HD float fix_to_f32(uint64_t raw, int frac) { return (float)raw * (1.0f / (float)(1u << frac)); }
// conversion32.rs:24-34
HD float scalar_to_f32(const sc &s, int n_bits, int frac) {
    uint64_t lo;
    if ((s.v[7] >> 24) != 0) { sc n; sc_neg(n, s); lo = ((uint64_t)n.v[1] << 32 | n.v[0]) & fix_max(n_bits); return -fix_to_f32(lo, frac); }
    lo = ((uint64_t)s.v[1] << 32 | s.v[0]) & fix_max(n_bits);
    return fix_to_f32(lo, frac);
}

// ---- generic bit-serial multiply by a 256-bit integer (table construction only) ---------------------------------
HDNI void ge_scalarmult_bits(ge_p3 &r, const uint32_t k[8], const ge_p3 &p) {
    ge_cached c; ge_p3_to_cached(c, p);
    ge_p3_0(r);
    int top = 255; while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        ge_p3_dbl(r, r);
        if ((k[i >> 5] >> (i & 31)) & 1) ge_add_cached(r, r, c);
    }
}

// ---- fixed-base tables: radix 256, signed digits; entry (w,k) = (k+1) * 256^w * P as affine niels ----------------
#define FB_WINDOWS 32
#define FB_ENTRIES 128
HDNI void fb_table_entry(ge_niels &n, const ge_p3 &p, int w, int k) {
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    s[w >> 2] = (uint32_t)(k + 1) << (8 * (w & 3));          // (k+1) <= 128 fits one byte lane
    ge_p3 q; ge_scalarmult_bits(q, s, p);
    fe zinv; fe_invert(zinv, q.Z);
    ge_p3_to_niels(n, q, zinv);
}
// r += k * P  using P's table; nwin = number of radix-256 windows to scan (32 for full scalars)
HDNI void fb_mul_acc(ge_p3 &r, const niels_st *tab, const sc &k, int nwin) {
    int16_t d[32]; sc_radix256(d, k);
    for (int i = 0; i < nwin; i++) {
        int di = d[i];
        if (di != 0) {
            ge_niels n; ld_niels(n, tab + i * FB_ENTRIES + (di > 0 ? di : -di) - 1);
            ge_madd_signed(r, r, n, di < 0);
        }
    }
}

// ---- variable-base ladders ----------------------------------------------------------------------------------------
// (a) NAF supplied by the caller (one scalar shared by a whole block: warp-uniform control flow).
//     w <= 5; table = odd multiples P, 3P, ..., (2^(w-1)-1)P in cached form.
HD void ge_scalarmult_naf(ge_p3 &r, const int8_t naf[256], const ge_p3 &p, int w) {
    ge_cached tab[8];
    const int nt = 1 << (w - 2);
    ge_p3 p2, cur = p;
    ge_p3_dbl(p2, p);
    ge_cached c2; ge_p3_to_cached(c2, p2);
    ge_p3_to_cached(tab[0], p);
    for (int i = 1; i < nt; i++) { ge_add_cached(cur, cur, c2); ge_p3_to_cached(tab[i], cur); }
    int top = 255; while (top >= 0 && naf[top] == 0) top--;
    ge_p3_0(r);
    for (int i = top; i >= 0; i--) {
        ge_p1p1 t; ge_dbl_p1p1(t, r.X, r.Y, r.Z); ge_dbl_fix(t);
        int d = naf[i];
        if (d == 0) {
            if (i == 0) ge_p1p1_to_p3(r, t);
            else { fe_mul(r.X, t.T, t.X); fe_mul(r.Y, t.Y, t.Z); fe_mul(r.Z, t.T, t.Z); }      // T not needed before a doubling
        } else {
            ge_p1p1_to_p3(r, t);
            if (d > 0) ge_add_p1p1(t, r, tab[d >> 1]); else ge_sub_p1p1(t, r, tab[(-d) >> 1]);
            if (i == 0) ge_p1p1_to_p3(r, t);
            else { fe_mul(r.X, t.T, t.X); fe_mul(r.Y, t.Y, t.Z); fe_mul(r.Z, t.T, t.Z); }
        }
    }
    if (top < 0) ge_p3_0(r);
}
// inner-product-argument generator fold: r = lo + s*hi, s given as a NAF
HD void ge_fold(ge_p3 &r, const int8_t naf[256], const ge_p3 &lo, const ge_p3 &hi, int w) {
    ge_p3 t; ge_scalarmult_naf(t, naf, hi, w);
    ge_add(r, t, lo);
}
// (b) per-lane scalar: signed radix-16, fixed schedule (4 doublings + 1 table add per digit), table 1P..8P
HD void sc_radix16(int8_t e[64], const sc &a) {
    for (int i = 0; i < 64; i++) e[i] = (int8_t)((a.v[i >> 3] >> (4 * (i & 7))) & 15);
    int carry = 0;
    for (int i = 0; i < 63; i++) { int v = e[i] + carry; carry = (v + 8) >> 4; e[i] = (int8_t)(v - (carry << 4)); }
    e[63] = (int8_t)(e[63] + carry);
}
struct ge_tab8 { ge_cached t[8]; };
HDNI void ge_tab8_build(ge_tab8 &tb, const ge_p3 &p) {
    ge_p3 cur = p; ge_p3_to_cached(tb.t[0], p);
    for (int i = 1; i < 8; i++) { ge_add_cached(cur, cur, tb.t[0]); ge_p3_to_cached(tb.t[i], cur); }
}
// r = a*P + b*Q with shared doublings (b/Q optional: pass nullptr tables to skip)
HDNI void ge_double_scalarmult_r16(ge_p3 &r, const sc &a, const ge_tab8 &tp, const sc *b, const ge_tab8 *tq) {
    int8_t ea[64], eb[64];
    sc_radix16(ea, a);
    if (b) sc_radix16(eb, *b);
    ge_p3_0(r);
    for (int i = 63; i >= 0; i--) {
        if (i != 63) {
            ge_p2 q; ge_p1p1 t;
            ge_dbl_p1p1(t, r.X, r.Y, r.Z); ge_dbl_fix(t); ge_p1p1_to_p2(q, t);
            ge_dbl_p1p1(t, q.X, q.Y, q.Z); ge_dbl_fix(t); ge_p1p1_to_p2(q, t);
            ge_dbl_p1p1(t, q.X, q.Y, q.Z); ge_dbl_fix(t); ge_p1p1_to_p2(q, t);
            ge_dbl_p1p1(t, q.X, q.Y, q.Z); ge_dbl_fix(t); ge_p1p1_to_p3(r, t);
        }
        int d = ea[i];
        if (d != 0) ge_add_cached_signed(r, r, tp.t[(d > 0 ? d : -d) - 1], d < 0);
        if (b) { d = eb[i]; if (d != 0) ge_add_cached_signed(r, r, tq->t[(d > 0 ? d : -d) - 1], d < 0); }
    }
}

// ---- BulletproofGens chain: SHAKE256("GeneratorsChain" || 'G'|'H' || u32le(party)) -------------------
struct gen_chain { sponge s; };
HD void gen_chain_init(gen_chain &c, int which, uint32_t party) {
    const uint8_t dom[15] = {'G', 'e', 'n', 'e', 'r', 'a', 't', 'o', 'r', 's', 'C', 'h', 'a', 'i', 'n'};
    uint8_t label[5] = {(uint8_t)which, (uint8_t)party, (uint8_t)(party >> 8), (uint8_t)(party >> 16), (uint8_t)(party >> 24)};
    sponge_init(c.s, 136); sponge_absorb(c.s, dom, 15); sponge_absorb(c.s, label, 5); sponge_finish(c.s, 0x1f);
}
HD void gen_chain_next(gen_chain &c, ge_p3 &p) { uint8_t u[64]; sponge_squeeze(c.s, u, 64); ge_from_uniform_bytes(p, u); }
HD void gen_chain_points(uint8_t *out, int which, uint32_t party, int n) {
    gen_chain c; gen_chain_init(c, which, party);
    for (int i = 0; i < n; i++) { ge_p3 p; gen_chain_next(c, p); ge_compress(out + 32 * i, p); }
}

// ---- nonce stream: draw k == ChaCha20 block k of `key`, wide-reduced -----------------------------------
HDNI void nonce_scalar(sc &out, const uint32_t key[8], uint64_t idx) { uint32_t w[16]; chacha20_block_words(w, key, idx); sc_from_wide_words(out, w); }

// ---- per-element square proof (square_proof/{mod,party,dealer}.rs; square_proof_vec/mod.rs:19-75,130-160) ----------------
HDNI void sq_transcript_challenge(sc &c, const uint8_t cl[32], const uint8_t csq[32], const uint8_t cpl[32], const uint8_t cpsq[32]) {
    transcript t; transcript_init(t, "SquareProof");                                       // square_proof_vec/mod.rs:52
    const uint8_t ds[19] = {'r', 'a', 'n', 'd', 'o', 'm', 'n', 'e', 's', 's', ' ', 'p', 'r', 'o', 'o', 'f', ' ', 'v', '1'};
    transcript_append(t, "dom-sep", ds, 19);                                               // rand_proof/transcript.rs:20-22
    transcript_append(t, "C_eg", cl, 32); transcript_append(t, "C_ped", csq, 32);          // dealer.rs:25-27
    transcript_append(t, "C_prime_eg", cpl, 32); transcript_append(t, "C_prime_ped", cpsq, 32);   // dealer.rs:49-53
    uint8_t buf[64]; transcript_challenge(t, "c", buf, 64);
    sc_from_bytes_wide(c, buf);                                                            // transcript.rs:40-44
}
// returns 0 ok, -4 if C_l does not decode, -1 NaN
HD int square_prove_one(uint8_t *proof160, uint8_t *commit64, float v, const uint8_t *cl32, const uint8_t *r1b, const uint8_t *r2b,
                        const uint32_t key[8], uint64_t idx, int n_bits, int frac, const niels_st *tabB, const niels_st *tabH) {
    sc m, r1, r2, msq, mp, r1p, r2p, c, zm, zr1, zr2, tmp;
    ge_p3 Cl, P;
    if (f32_to_scalar(m, v, n_bits, frac)) return -1;
    if (!ge_decompress(Cl, cl32)) return -4;
    sc_from_bytes_mod_order(r1, r1b); sc_from_bytes_mod_order(r2, r2b);
    sc_mul(msq, m, m);
    ge_p3_0(P); fb_mul_acc(P, tabB, msq, 32); fb_mul_acc(P, tabH, r2, 32);                  // c_sq  (party.rs:34-35)
    for (int i = 0; i < 32; i++) commit64[i] = cl32[i];
    ge_compress(commit64 + 32, P);
    nonce_scalar(mp, key, 3 * idx); nonce_scalar(r1p, key, 3 * idx + 1); nonce_scalar(r2p, key, 3 * idx + 2);   // party.rs:40-42
    ge_p3_0(P); fb_mul_acc(P, tabB, mp, 32); fb_mul_acc(P, tabH, r1p, 32);                  // c'     (:44)
    ge_compress(proof160, P);
    ge_tab8 tb; ge_tab8_build(tb, Cl);
    ge_double_scalarmult_r16(P, mp, tb, nullptr, nullptr); fb_mul_acc(P, tabH, r2p, 32);    // c_sq' = m' C_l + r2' H (:46-50)
    ge_compress(proof160 + 32, P);
    sq_transcript_challenge(c, commit64, commit64 + 32, proof160, proof160 + 32);
    sc_muladd(zm, m, c, mp); sc_muladd(zr1, r1, c, r1p);                                   // party.rs:146-152
    sc_mul(tmp, m, r1); sc_sub(tmp, r2, tmp); sc_muladd(zr2, tmp, c, r2p);
    sc_tobytes(proof160 + 64, zm); sc_tobytes(proof160 + 96, zr1); sc_tobytes(proof160 + 128, zr2);
    return 0;
}
// returns 1 valid, 0 invalid, -1 FormatError (square_proof/mod.rs:77-112,127-146)
HD int square_verify_one(const uint8_t *proof160, const uint8_t *commit64, const niels_st *tabB, const niels_st *tabH) {
    ge_p3 Cl, Csq, Cp, Csqp, lhs, rhs, T;
    sc zm, zr1, zr2, c, nc;
    bool ok = ge_decompress(Cl, commit64) & ge_decompress(Csq, commit64 + 32) & ge_decompress(Cp, proof160) & ge_decompress(Csqp, proof160 + 32);
    sc_frombytes(zm, proof160 + 64); sc_frombytes(zr1, proof160 + 96); sc_frombytes(zr2, proof160 + 128);
    ok = ok && sc_is_canonical(zm) && sc_is_canonical(zr1) && sc_is_canonical(zr2);
    if (!ok) return -1;
    sq_transcript_challenge(c, commit64, commit64 + 32, proof160, proof160 + 32);
    // z_m B + z_r1 H == C'_l + c C_l
    ge_tab8 tl, ts; ge_tab8_build(tl, Cl); ge_tab8_build(ts, Csq);
    ge_p3_0(lhs); fb_mul_acc(lhs, tabB, zm, 32); fb_mul_acc(lhs, tabH, zr1, 32);
    ge_double_scalarmult_r16(T, c, tl, nullptr, nullptr); ge_add(rhs, Cp, T);
    bool v1 = ge_eq(lhs, rhs);
    // z_m C_l + z_r2 H == C'_sq + c C_sq   <=>   z_m C_l - c C_sq + z_r2 H == C'_sq
    sc_neg(nc, c);
    ge_double_scalarmult_r16(lhs, zm, tl, &nc, &ts); fb_mul_acc(lhs, tabH, zr2, 32);
    bool v2 = ge_eq(lhs, Csqp);
    return (v1 && v2) ? 1 : 0;
}

// ---- per-element ElGamal randomness proof (rand_proof/{mod,party,dealer}.rs; rand_proof_vec/mod.rs:14-118) -- enc type 2 -----------------
//   pair = (L, R) = (m B + r H, r B) (L may be an existing commitment: PartyExisting), C' = commit(m', r'), c = H(pair, C'),
//   z_m = m' + m c, z_r = r' + r c;  proof = C'_L | C'_R | z_m | z_r (128 B).  Nonces: blocks 2i, 2i+1 of the key stream.
HDNI void rp_transcript_challenge(sc &c, const uint8_t pair[64], const uint8_t cprime[64]) {
    transcript t; transcript_init(t, "RandProof");                                           // rand_proof_vec/mod.rs:32,73,104
    const uint8_t ds[19] = {'r', 'a', 'n', 'd', 'o', 'm', 'n', 'e', 's', 's', ' ', 'p', 'r', 'o', 'o', 'f', ' ', 'v', '1'};
    transcript_append(t, "dom-sep", ds, 19);
    transcript_append(t, "C", pair, 64); transcript_append(t, "C_prime", cprime, 64);       // dealer.rs:20-21,42
    uint8_t buf[64]; transcript_challenge(t, "c", buf, 64);
    sc_from_bytes_wide(c, buf);
}
// value_com32 null: L = commit(m, r) (Party::new), else the existing commitment.  returns 0, -1 NaN, -4 existing commitment does not decode
HD int rand_prove_one(uint8_t *proof128, uint8_t *pair64, float v, const uint8_t *value_com32, const uint8_t *rb, const uint32_t key[8], uint64_t idx,
                      int n_bits, int frac, const niels_st *tabB, const niels_st *tabH) {
    sc m, r, mp, rp, c, zm, zr; ge_p3 P;
    if (f32_to_scalar(m, v, n_bits, frac)) return -1;
    sc_from_bytes_mod_order(r, rb);
    if (value_com32) { if (!ge_decompress(P, value_com32)) return -4; for (int i = 0; i < 32; i++) pair64[i] = value_com32[i]; }
    else { ge_p3_0(P); fb_mul_acc(P, tabB, m, 32); fb_mul_acc(P, tabH, r, 32); ge_compress(pair64, P); }
    ge_p3_0(P); fb_mul_acc(P, tabB, r, 32); ge_compress(pair64 + 32, P);                     // R = r B (el_gamal.rs:57-69)
    nonce_scalar(mp, key, 2 * idx); nonce_scalar(rp, key, 2 * idx + 1);                      // party.rs:23-24
    ge_p3_0(P); fb_mul_acc(P, tabB, mp, 32); fb_mul_acc(P, tabH, rp, 32); ge_compress(proof128, P);
    ge_p3_0(P); fb_mul_acc(P, tabB, rp, 32); ge_compress(proof128 + 32, P);
    rp_transcript_challenge(c, pair64, proof128);
    sc_muladd(zm, m, c, mp); sc_muladd(zr, r, c, rp);                                       // party.rs:76-80
    sc_tobytes(proof128 + 64, zm); sc_tobytes(proof128 + 96, zr);
    return 0;
}
// 1 valid, 0 invalid, -1 FormatError (rand_proof/mod.rs:64-85,99-118)
HD int rand_verify_one(const uint8_t *proof128, const uint8_t *pair64, const niels_st *tabB, const niels_st *tabH) {
    ge_p3 L, R, Lp, Rp, lhs, rhs, T; sc zm, zr, c;
    bool ok = ge_decompress(L, pair64) & ge_decompress(R, pair64 + 32) & ge_decompress(Lp, proof128) & ge_decompress(Rp, proof128 + 32);
    sc_frombytes(zm, proof128 + 64); sc_frombytes(zr, proof128 + 96);
    ok = ok && sc_is_canonical(zm) && sc_is_canonical(zr);
    if (!ok) return -1;
    rp_transcript_challenge(c, pair64, proof128);
    ge_tab8 tl, tr; ge_tab8_build(tl, L); ge_tab8_build(tr, R);
    ge_p3_0(lhs); fb_mul_acc(lhs, tabB, zm, 32); fb_mul_acc(lhs, tabH, zr, 32);             // z_m B + z_r H == C'_L + c L
    ge_double_scalarmult_r16(T, c, tl, nullptr, nullptr); ge_add(rhs, Lp, T);
    const bool v1 = ge_eq(lhs, rhs);
    ge_p3_0(lhs); fb_mul_acc(lhs, tabB, zr, 32);                                             // z_r B == C'_R + c R
    ge_double_scalarmult_r16(T, c, tr, nullptr, nullptr); ge_add(rhs, Rp, T);
    return (v1 && ge_eq(lhs, rhs)) ? 1 : 0;
}

// ---- per-element square + randomness proof (square_rand_proof/*; square_rand_proof_vec/mod.rs:18-160) -- enc type 3 ------------------------
//   commitments = c.L | c.R | c_sq (96 B), proof = C'.L | C'.R | C'_sq | z_m | z_r1 | z_r2 (192 B).  Nonces: blocks 3i, 3i+1, 3i+2.
HDNI void srp_transcript_challenge(sc &c, const uint8_t com[96], const uint8_t cp[96]) {
    transcript t; transcript_init(t, "SquareRandProof");                                     // square_rand_proof_vec/mod.rs:51,105,142
    const uint8_t ds[19] = {'r', 'a', 'n', 'd', 'o', 'm', 'n', 'e', 's', 's', ' ', 'p', 'r', 'o', 'o', 'f', ' ', 'v', '1'};
    transcript_append(t, "dom-sep", ds, 19);
    transcript_append(t, "C_eg", com, 64); transcript_append(t, "C_ped", com + 64, 32);     // dealer.rs:23-25
    transcript_append(t, "C_prime_eg", cp, 64); transcript_append(t, "C_prime_ped", cp + 64, 32);
    uint8_t buf[64]; transcript_challenge(t, "c", buf, 64);
    sc_from_bytes_wide(c, buf);
}
HD int square_rand_prove_one(uint8_t *proof192, uint8_t *com96, float v, const uint8_t *value_com32, const uint8_t *r1b, const uint8_t *r2b, const uint32_t key[8], uint64_t idx,
                             int n_bits, int frac, const niels_st *tabB, const niels_st *tabH) {
    sc m, r1, r2, msq, mp, r1p, r2p, c, zm, zr1, zr2, tmp; ge_p3 Cl, P;
    if (f32_to_scalar(m, v, n_bits, frac)) return -1;
    sc_from_bytes_mod_order(r1, r1b); sc_from_bytes_mod_order(r2, r2b);
    if (value_com32) { if (!ge_decompress(Cl, value_com32)) return -4; for (int i = 0; i < 32; i++) com96[i] = value_com32[i]; }
    else { ge_p3_0(Cl); fb_mul_acc(Cl, tabB, m, 32); fb_mul_acc(Cl, tabH, r1, 32); ge_compress(com96, Cl); }
    ge_p3_0(P); fb_mul_acc(P, tabB, r1, 32); ge_compress(com96 + 32, P);                     // c.R = r1 B
    sc_mul(msq, m, m); ge_p3_0(P); fb_mul_acc(P, tabB, msq, 32); fb_mul_acc(P, tabH, r2, 32); ge_compress(com96 + 64, P);      // c_sq (party.rs:31-32)
    nonce_scalar(mp, key, 3 * idx); nonce_scalar(r1p, key, 3 * idx + 1); nonce_scalar(r2p, key, 3 * idx + 2);
    ge_p3_0(P); fb_mul_acc(P, tabB, mp, 32); fb_mul_acc(P, tabH, r1p, 32); ge_compress(proof192, P);      // C'.L
    ge_p3_0(P); fb_mul_acc(P, tabB, r1p, 32); ge_compress(proof192 + 32, P);                              // C'.R
    ge_tab8 tb; ge_tab8_build(tb, Cl);
    ge_double_scalarmult_r16(P, mp, tb, nullptr, nullptr); fb_mul_acc(P, tabH, r2p, 32); ge_compress(proof192 + 64, P);       // C'_sq = m' c.L + r2' H
    srp_transcript_challenge(c, com96, proof192);
    sc_muladd(zm, m, c, mp); sc_muladd(zr1, r1, c, r1p);
    sc_mul(tmp, m, r1); sc_sub(tmp, r2, tmp); sc_muladd(zr2, tmp, c, r2p);
    sc_tobytes(proof192 + 96, zm); sc_tobytes(proof192 + 128, zr1); sc_tobytes(proof192 + 160, zr2);
    return 0;
}
HD int square_rand_verify_one(const uint8_t *proof192, const uint8_t *com96, const niels_st *tabB, const niels_st *tabH) {
    ge_p3 L, R, Csq, Lp, Rp, Csqp, lhs, rhs, T; sc zm, zr1, zr2, c, nc;
    bool ok = ge_decompress(L, com96) & ge_decompress(R, com96 + 32) & ge_decompress(Csq, com96 + 64) & ge_decompress(Lp, proof192) & ge_decompress(Rp, proof192 + 32) & ge_decompress(Csqp, proof192 + 64);
    sc_frombytes(zm, proof192 + 96); sc_frombytes(zr1, proof192 + 128); sc_frombytes(zr2, proof192 + 160);
    ok = ok && sc_is_canonical(zm) && sc_is_canonical(zr1) && sc_is_canonical(zr2);
    if (!ok) return -1;
    srp_transcript_challenge(c, com96, proof192);
    ge_tab8 tl, tr, ts; ge_tab8_build(tl, L); ge_tab8_build(tr, R); ge_tab8_build(ts, Csq);
    ge_p3_0(lhs); fb_mul_acc(lhs, tabB, zm, 32); fb_mul_acc(lhs, tabH, zr1, 32);            // mod.rs:94-99
    ge_double_scalarmult_r16(T, c, tl, nullptr, nullptr); ge_add(rhs, Lp, T);
    bool v = ge_eq(lhs, rhs);
    ge_p3_0(lhs); fb_mul_acc(lhs, tabB, zr1, 32);
    ge_double_scalarmult_r16(T, c, tr, nullptr, nullptr); ge_add(rhs, Rp, T);
    v = v && ge_eq(lhs, rhs);
    sc_neg(nc, c);                                                                            // z_m c.L + z_r2 H == C'_sq + c c_sq (mod.rs:101-109)
    ge_double_scalarmult_r16(lhs, zm, tl, &nc, &ts); fb_mul_acc(lhs, tabH, zr2, 32);
    v = v && ge_eq(lhs, Csqp);
    return v ? 1 : 0;
}
