// TEST TOOL: compiles the product's __host__ __device__ math headers (rofl-project-code_b200/csrc/*.cuh)
// for the CPU with FE_CHECK_BOUNDS so that the exact device arithmetic can be compared with the oracle in
// this GPU-less container.  Not part of the product and never used as a fallback: the product library only
// exposes CUDA entry points.
// (the saturated-limb field has no limb-scale preconditions to assert)
#include "../../rofl-project-code_b200/csrc/ge25519.cuh"
#include "../../rofl-project-code_b200/csrc/hash.cuh"
#include "../../rofl-project-code_b200/csrc/devfn.cuh"
#include <vector>
#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT void hs_fe_mul(uint8_t *o, const uint8_t *a, const uint8_t *b) { fe x, y, r; fe_frombytes(x, a); fe_frombytes(y, b); fe_mul(r, x, y); fe_tobytes(o, r); }
EXPORT void hs_fe_sq(uint8_t *o, const uint8_t *a) { fe x, r; fe_frombytes(x, a); fe_sq(r, x); fe_tobytes(o, r); }
EXPORT void hs_fe_invert(uint8_t *o, const uint8_t *a) { fe x, r; fe_frombytes(x, a); fe_invert(r, x); fe_tobytes(o, r); }
// (a+b)*(c-d) with unreduced operands, to exercise the scale rules
EXPORT void hs_fe_mix(uint8_t *o, const uint8_t *a, const uint8_t *b, const uint8_t *c, const uint8_t *d) {
    fe x, y, z, w, s, t, r; fe_frombytes(x, a); fe_frombytes(y, b); fe_frombytes(z, c); fe_frombytes(w, d);
    fe_add(s, x, y); fe_sub(t, z, w); fe_mul(r, t, s); fe_tobytes(o, r);
}
EXPORT void hs_sc_mul(uint8_t *o, const uint8_t *a, const uint8_t *b) { sc x, y, r; sc_frombytes(x, a); sc_frombytes(y, b); sc_mul(r, x, y); sc_tobytes(o, r); }
EXPORT void hs_sc_add(uint8_t *o, const uint8_t *a, const uint8_t *b) { sc x, y, r; sc_frombytes(x, a); sc_frombytes(y, b); sc_add(r, x, y); sc_tobytes(o, r); }
EXPORT void hs_sc_sub(uint8_t *o, const uint8_t *a, const uint8_t *b) { sc x, y, r; sc_frombytes(x, a); sc_frombytes(y, b); sc_sub(r, x, y); sc_tobytes(o, r); }
EXPORT void hs_sc_wide(uint8_t *o, const uint8_t *a) { sc r; sc_from_bytes_wide(r, a); sc_tobytes(o, r); }
EXPORT void hs_sc_invert(uint8_t *o, const uint8_t *a) { sc x, r; sc_frombytes(x, a); sc_invert(r, x); sc_tobytes(o, r); }
EXPORT void hs_sc_invert_vartime(uint8_t *o, const uint8_t *a) { sc x, r; sc_frombytes(x, a); sc_invert_vartime(r, x); sc_tobytes(o, r); }
EXPORT int hs_decompress_compress(uint8_t *o, const uint8_t *a) { ge_p3 p; bool ok = ge_decompress(p, a); if (ok) ge_compress(o, p); return ok; }
EXPORT void hs_from_uniform(uint8_t *o, const uint8_t *a) { ge_p3 p; ge_from_uniform_bytes(p, a); ge_compress(o, p); }
EXPORT int hs_point_add(uint8_t *o, const uint8_t *a, const uint8_t *b, int sub) {
    ge_p3 p, q, r; if (!ge_decompress(p, a) || !ge_decompress(q, b)) return 0;
    if (sub) ge_sub(r, p, q); else ge_add(r, p, q); ge_compress(o, r); return 1;
}
EXPORT int hs_point_dbl(uint8_t *o, const uint8_t *a, int n) { ge_p3 p; if (!ge_decompress(p, a)) return 0; for (int i = 0; i < n; i++) ge_p3_dbl(p, p); ge_compress(o, p); return 1; }
EXPORT int hs_point_eq(const uint8_t *a, const uint8_t *b) { ge_p3 p, q; ge_decompress(p, a); ge_decompress(q, b); return ge_eq(p, q); }
// variable-base multiply through the width-w NAF ladder used by the fold kernel
EXPORT int hs_scalarmult_naf(uint8_t *o, const uint8_t *s, const uint8_t *a, int w) {
    ge_p3 p, r; sc k; if (!ge_decompress(p, a)) return 0; sc_frombytes(k, s);
    int8_t naf[256]; sc_naf(naf, k, w);
    ge_scalarmult_naf(r, naf, p, w); ge_compress(o, r); return 1;
}
// r = lo + s*hi (fold step)
EXPORT int hs_fold(uint8_t *o, const uint8_t *s, const uint8_t *lo, const uint8_t *hi, int w) {
    ge_p3 pl, ph, r; sc k; if (!ge_decompress(pl, lo) || !ge_decompress(ph, hi)) return 0; sc_frombytes(k, s);
    int8_t naf[256]; sc_naf(naf, k, w);
    ge_fold(r, naf, pl, ph, w); ge_compress(o, r); return 1;
}
// fixed-base: build a radix-256 table for point a, then multiply
static std::vector<niels_st> g_tab;
EXPORT int hs_fixed_table(const uint8_t *a) {
    ge_p3 p; if (!ge_decompress(p, a)) return 0;
    g_tab.resize(FB_WINDOWS * FB_ENTRIES);
    for (int w = 0; w < FB_WINDOWS; w++) for (int k = 0; k < FB_ENTRIES; k++) { ge_niels e; fb_table_entry(e, p, w, k); st_niels(&g_tab[w * FB_ENTRIES + k], e); }
    return 1;
}
EXPORT void hs_fixed_mul(uint8_t *o, const uint8_t *s) { sc k; sc_frombytes(k, s); ge_p3 r; ge_p3_0(r); fb_mul_acc(r, g_tab.data(), k, 32); ge_compress(o, r); }
EXPORT void hs_sha3_512(uint8_t *o, const uint8_t *in, size_t n) { sha3_512(o, in, n); }
EXPORT void hs_shake256(uint8_t *o, size_t on, const uint8_t *in, size_t n) { sponge s; sponge_init(s, 136); sponge_absorb(s, in, n); sponge_finish(s, 0x1f); sponge_squeeze(s, o, on); }
EXPORT void hs_merlin_simple(uint8_t *o, size_t on, const char *proto, const char *label, const uint8_t *msg, size_t n, const char *clabel) {
    transcript t; transcript_init(t, proto); transcript_append(t, label, msg, n); transcript_challenge(t, clabel, o, on);
}
EXPORT void hs_chacha(uint8_t *o, const uint8_t *key, uint64_t ctr) { uint32_t k[8], w[16]; memcpy(k, key, 32); chacha20_block_words(w, k, ctr); memcpy(o, w, 64); }
EXPORT void hs_gen_chain(uint8_t *o, int which, uint32_t party, int n) {
    gen_chain_points(o, which, party, n);
}
EXPORT int hs_f32_to_raw(uint64_t *raw, float x, int n_bits, int frac) { return f32_to_fix(raw, x, n_bits, frac); }
EXPORT void hs_square_prove_one(uint8_t *proof160, uint8_t *commit64, float v, const uint8_t *cl32, const uint8_t *r1, const uint8_t *r2,
                                const uint8_t *key, uint64_t idx, int n_bits, int frac, const uint8_t *Bp, const uint8_t *Hp) {
    std::vector<niels_st> tb(FB_WINDOWS * FB_ENTRIES), th(FB_WINDOWS * FB_ENTRIES);
    ge_p3 B, H; ge_decompress(B, Bp); ge_decompress(H, Hp);
    for (int w = 0; w < FB_WINDOWS; w++) for (int k = 0; k < FB_ENTRIES; k++) { ge_niels e; fb_table_entry(e, B, w, k); st_niels(&tb[w * FB_ENTRIES + k], e); fb_table_entry(e, H, w, k); st_niels(&th[w * FB_ENTRIES + k], e); }
    uint32_t kw[8]; memcpy(kw, key, 32);
    square_prove_one(proof160, commit64, v, cl32, r1, r2, kw, idx, n_bits, frac, tb.data(), th.data());
}
EXPORT int hs_square_verify_one(const uint8_t *proof160, const uint8_t *commit64, const uint8_t *Bp, const uint8_t *Hp) {
    std::vector<niels_st> tb(FB_WINDOWS * FB_ENTRIES), th(FB_WINDOWS * FB_ENTRIES);
    ge_p3 B, H; ge_decompress(B, Bp); ge_decompress(H, Hp);
    for (int w = 0; w < FB_WINDOWS; w++) for (int k = 0; k < FB_ENTRIES; k++) { ge_niels e; fb_table_entry(e, B, w, k); st_niels(&tb[w * FB_ENTRIES + k], e); fb_table_entry(e, H, w, k); st_niels(&th[w * FB_ENTRIES + k], e); }
    return square_verify_one(proof160, commit64, tb.data(), th.data());
}
