// librofl_b200_bindings32.so -- the reference's own `extern "C"` ABI (rofl_crypto/src/bindings32.rs:29-764), symbol for symbol, on top of
// librofl_b200.so.  A ctypes consumer of the reference's `librofl_crypto.so` (crate-type dylib, rofl_crypto/Cargo.toml:10) loads this file
// instead and calls the same names with the same arguments; every heavy operation runs on the GPU through include/rofl_b200.h.
//
//   #[repr(C)] PyVec { data, len }, PyRes { ret, msg, res }            bindings32.rs:29-40
//   ownership: inputs are borrowed; outputs are heap blocks that are NEVER freed by this library (the reference mem::forget()s them,
//   bindings32.rs:795-802); rofl_bindings32_free() is an addition for consumers that want to give them back.
//   errors: ret = 1 + message (bindings32.rs:814-823); null inputs abort like the reference's assert!s.
//   payloads: bincode 1.3 (serde_vec.rs:9-71) --  Vec<T> = u64-LE count + elements;  Scalar / RistrettoPoint = 32 raw bytes (serde tuple of 32
//   u8 in curve25519-dalek-ng 4.1 [UPSTREAM-RECALL: the crate source is not under /root/reference; no test of the reference pins the layout]);
//   RangeProof / ElGamalPair / RandProof / SquareRandProof / SquareRandProofCommitments = serialize_bytes(to_bytes()) = u64-LE length + bytes
//   (rand_proof/el_gamal.rs:197-204, rand_proof/mod.rs:118-125, square_rand_proof/mod.rs:149-156, square_rand_proof/pedersen.rs:49-56).
//   fixed point: the reference bakes N_BITS / frac / PRECOMP_BIAS in with cargo features (fp.rs:35-137); here rofl_bindings32_configure(), or the
//   environment (ROFL_FP_BITS, ROFL_FP_FRAC, ROFL_B200_DEVICE); default = the reference's default features (16 / 7, PRECOMP_BIAS 8 -> table 2^16).
//   randomness: every nonce comes from a fresh OS-random seed per call (the reference uses thread_rng).
#include "../../include/rofl_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

extern "C" {
struct PyVec { const void *data; size_t len; };
struct PyRes { size_t ret; const char *msg; const void *res; };
}

namespace {
typedef std::vector<uint8_t> bytes;
int g_bits = 16, g_frac = 7, g_device = 0, g_bias = 8; bool g_cfg = false;
rofl_ctx *g_ctx = nullptr; std::mutex g_mu;
int bias_of(int bits) { return bits == 8 ? 3 : bits == 16 ? 7 : bits == 32 ? 7 : 0; }           // fp.rs:45,61,77,98 (the featureless default is 8, fp.rs:133)
void cfg_from_env() {
    if (g_cfg) return;
    if (const char *v = getenv("ROFL_FP_BITS")) { g_bits = atoi(v); g_bias = bias_of(g_bits); }
    if (const char *v = getenv("ROFL_FP_FRAC")) g_frac = atoi(v);
    if (const char *v = getenv("ROFL_B200_DEVICE")) g_device = atoi(v);
    g_cfg = true;
}
rofl_ctx *ctx() {
    std::lock_guard<std::mutex> lk(g_mu);
    cfg_from_env();
    if (!g_ctx && rofl_ctx_create(&g_ctx, g_device) != 0) { fprintf(stderr, "rofl_b200 bindings32: %s\n", rofl_last_error()); abort(); }      // no CPU fallback
    return g_ctx;
}
int bsgs_bits() { return g_bits <= 16 ? g_bits : 16; }                                           // fp.rs:47,63,84,101
uint64_t default_table() { return (uint64_t)1 << (bsgs_bits() / 2 + g_bias); }                  // BSGSTable::default, bsgs32.rs:36-38
void need(const void *p) { if (!p) { fprintf(stderr, "rofl_b200 bindings32: null argument\n"); abort(); } }      // assert!(!ptr.is_null())
void fresh_seed(uint8_t s[32]) { FILE *f = fopen("/dev/urandom", "rb"); if (!f || fread(s, 1, 32, f) != 32) abort(); fclose(f); }

// ---- bincode ------------------------------------------------------------------------------------------------------------------------------
uint64_t rd64(const uint8_t *p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
void wr64(bytes &o, uint64_t v) { for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i))); }
// Vec<Scalar> / Vec<RistrettoPoint>: u64 count + count x 32 bytes
bool de_vec32(const uint8_t *p, size_t len, bytes &out) {
    if (len < 8) return false; const uint64_t n = rd64(p);
    if (len < 8 + 32 * n) return false;
    out.assign(p + 8, p + 8 + 32 * n); return true;
}
bytes ser_vec32(const uint8_t *items, size_t n) { bytes o; o.reserve(8 + 32 * n); wr64(o, n); o.insert(o.end(), items, items + 32 * n); return o; }
// Vec<T>, T = serialize_bytes(fixed width w): u64 count + count x (u64 w + w bytes)
bool de_vec_bytes(const uint8_t *p, size_t len, size_t w, bytes &out) {
    if (len < 8) return false; const uint64_t n = rd64(p); size_t off = 8; out.clear(); out.reserve(n * w);
    for (uint64_t i = 0; i < n; i++) { if (off + 8 > len || rd64(p + off) != w || off + 8 + w > len) return false; out.insert(out.end(), p + off + 8, p + off + 8 + w); off += 8 + w; }
    return true;
}
bytes ser_vec_bytes(const uint8_t *items, size_t n, size_t w) { bytes o; o.reserve(8 + n * (8 + w)); wr64(o, n); for (size_t i = 0; i < n; i++) { wr64(o, w); o.insert(o.end(), items + i * w, items + (i + 1) * w); } return o; }
// Vec<RangeProof>: proofs of one call all have the same length
bool de_range_proofs(const uint8_t *p, size_t len, bytes &out, size_t &plen, size_t &n) {
    if (len < 8) return false; n = rd64(p); size_t off = 8; out.clear(); plen = 0;
    for (uint64_t i = 0; i < n; i++) { if (off + 8 > len) return false; const uint64_t w = rd64(p + off); if (i == 0) plen = w; if (w != plen || off + 8 + w > len) return false; out.insert(out.end(), p + off + 8, p + off + 8 + w); off += 8 + w; }
    return true;
}

// ---- leaked outputs (bindings32.rs:795-823) -----------------------------------------------------------------------------------------------
template <class T> PyVec leak(const std::vector<T> &v) {
    T *p = (T *)malloc(v.size() * sizeof(T) + 1); if (v.size()) memcpy(p, v.data(), v.size() * sizeof(T));
    return PyVec{p, v.size()};
}
PyVec leak_pyvecs(const std::vector<PyVec> &v) { return leak(v); }
PyRes ok_res(const void *p) { return PyRes{0, nullptr, p}; }
PyRes ok_bool(bool b) { uint8_t *p = (uint8_t *)malloc(1); *p = b ? 1 : 0; return ok_res(p); }
PyRes ok_pyvecs(const std::vector<PyVec> &v) { PyVec *box = (PyVec *)malloc(sizeof(PyVec)); *box = leak_pyvecs(v); return ok_res(box); }
PyRes err_res(const std::string &m) { char *s = (char *)malloc(m.size() + 1); memcpy(s, m.c_str(), m.size() + 1); return PyRes{1, s, nullptr}; }
// Display strings of the reference's error enums (range_proof_vec/errors.rs:5-19, l2_range_proof_vec/errors.rs:4-31, bulletproofs ProofError)
std::string err_text(int rc) {
    switch (rc) {
    case ROFL_ERR_VALUE_OUT_OF_RANGE: return "Fixed precision representation does exceed prove range bounds";
    case ROFL_ERR_OVERFLOW: return "The scalar calculation does not match the floating point calculation";
    case ROFL_ERR_NORM_OUT_OF_RANGE: return "The scalar L2 norm exceeds prove range bounds";
    case ROFL_ERR_FORMAT: return "Internal error during proof creation: Proof data could not be parsed.";
    case ROFL_ERR_BITSIZE: return "Internal error during proof creation: Invalid bitsize, must have n = 8,16,32,64.";
    case ROFL_ERR_GENS: return "Invalid generators size, too few generators for proof";
    case ROFL_ERR_POINT: return "A commitment does not decode";
    default: return std::string("rofl_b200 error ") + std::to_string(rc) + ": " + rofl_last_error();
    }
}
}  // namespace

extern "C" {
// ---- additions (not in the reference) -------------------------------------------------------------------------------------------------------
void rofl_bindings32_configure(int n_bits, int frac, int device) { std::lock_guard<std::mutex> lk(g_mu); g_bits = n_bits; g_frac = frac; g_device = device; g_bias = bias_of(n_bits); g_cfg = true; }
void rofl_bindings32_free(void *p) { free(p); }

// ---- bindings32.rs:42-60 ---------------------------------------------------------------------------------------------------------------------
PyVec say_hello() {
    printf("Hello world\n");
    static const uint8_t x[32] = {0x4e, 0x5a, 0xb4, 0x34, 0x5d, 0x47, 0x08, 0x84, 0x59, 0x13, 0xb4, 0x64, 0x1b, 0xc2, 0x7d, 0x52, 0x52, 0xa5, 0x85, 0x10, 0x1b, 0xcc, 0x42, 0x44, 0xd4, 0x49, 0xf4, 0xa8, 0x79, 0xd9, 0xf2, 0x04};
    return leak(ser_vec32(x, 1));                                     // create_x (:825-833) is canonical already
}
// ---- :64-86 add_commitments: element-wise sum of `len` serialized commitment vectors -----------------------------------------------------------
PyVec add_commitments(const uint8_t *const *ptr_arr, const size_t *len_arr, size_t len) {
    need(ptr_arr); need(len_arr);
    bytes all; size_t D = 0;
    for (size_t k = 0; k < len; k++) { bytes v; if (!de_vec32(ptr_arr[k], len_arr[k], v)) abort(); if (k == 0) D = v.size() / 32; if (v.size() / 32 != D) abort(); all.insert(all.end(), v.begin(), v.end()); }
    bytes out(32 * D);
    if (D && rofl_aggregate(ctx(), all.data(), len, D, 0, out.data()) != 0) abort();      // add_rp_vec_vec starts from the identity (pedersen_ops.rs:61-74)
    return leak(ser_vec32(out.data(), D));
}
// ---- :90-114 add_commitments_transposed: the sum of EACH vector, one bincode RistrettoPoint per input ------------------------------------------
PyVec add_commitments_transposed(const uint8_t *const *ptr_arr, const size_t *len_arr, size_t len) {
    need(ptr_arr); need(len_arr);
    std::vector<PyVec> outs;
    for (size_t k = 0; k < len; k++) {
        bytes v; if (!de_vec32(ptr_arr[k], len_arr[k], v) || v.empty()) abort();          // reduce().unwrap() panics on an empty vector
        bytes s(32); if (rofl_aggregate(ctx(), v.data(), v.size() / 32, 1, 0, s.data()) != 0) abort();      // n "clients" x 1 point
        outs.push_back(leak(s));
    }
    return leak_pyvecs(outs);
}
// ---- :118-149 commitments ------------------------------------------------------------------------------------------------------------------------
PyVec commit_no_blinding(const float *value_ptr, size_t len) {
    need(value_ptr); bytes L(32 * len);
    if (len && rofl_commit(ctx(), value_ptr, nullptr, len, g_bits, g_frac, L.data(), nullptr) != 0) abort();
    return leak(ser_vec32(L.data(), len));
}
PyVec commit(const float *value_ptr, size_t value_len, const uint8_t *blinding_ptr, size_t blinding_len) {
    need(value_ptr); need(blinding_ptr);
    bytes bl; if (!de_vec32(blinding_ptr, blinding_len, bl) || bl.size() / 32 != value_len) abort();
    bytes L(32 * value_len);
    if (value_len && rofl_commit(ctx(), value_ptr, bl.data(), value_len, g_bits, g_frac, L.data(), nullptr) != 0) abort();
    return leak(ser_vec32(L.data(), value_len));
}
// ---- :154-167 generate_cancelling_blindings (pedersen_ops.rs:110-122): n_vec random vectors whose element-wise sum is zero -------------------------
PyVec generate_cancelling_blindings(size_t n_vec, size_t n_dim) {
    if (n_vec == 0) abort();                                         // `0..n_vec - 1` underflows in the reference
    std::vector<bytes> vs(n_vec, bytes(32 * n_dim)); bytes sum(32 * n_dim, 0);
    for (size_t k = 0; k < n_vec; k++) { uint8_t seed[32]; fresh_seed(seed); rofl_rnd_scalar_vec(seed, n_dim, vs[k].data()); }
    for (size_t k = 0; k + 1 < n_vec; k++) rofl_scalar_ops(0, sum.data(), vs[k].data(), n_dim, sum.data());
    rofl_scalar_ops(1, sum.data(), nullptr, n_dim, vs[n_vec - 1].data());
    std::vector<PyVec> outs; for (auto &v : vs) outs.push_back(leak(ser_vec32(v.data(), n_dim)));
    return leak_pyvecs(outs);
}
// ---- :169-210 selectors --------------------------------------------------------------------------------------------------------------------------
static PyVec select32(const uint8_t *p, size_t len, const size_t *idx, size_t n) {
    bytes v; if (!de_vec32(p, len, v)) abort(); bytes o(32 * n);
    for (size_t i = 0; i < n; i++) { if (idx[i] >= v.size() / 32) abort(); memcpy(&o[32 * i], &v[32 * idx[i]], 32); }
    return leak(ser_vec32(o.data(), n));
}
PyVec select_blindings(const uint8_t *blinding_ptr, size_t blinding_len, const size_t *indices_ptr, size_t indices_len) { return select32(blinding_ptr, blinding_len, indices_ptr, indices_len); }
PyVec select_commitments(const uint8_t *commit_ptr, size_t commit_len, const size_t *indices_ptr, size_t indices_len) { return select32(commit_ptr, commit_len, indices_ptr, indices_len); }
// ---- :213-221 extract_values: default_discrete_log_vec + scalar_to_f32_vec -> PyVec of f32 ----------------------------------------------------------
PyVec extract_values(const uint8_t *bytes_ptr, size_t len) {
    need(bytes_ptr); bytes v; if (!de_vec32(bytes_ptr, len, v)) abort();
    const size_t D = v.size() / 32; std::vector<float> f(D);
    if (D && rofl_dlog(ctx(), v.data(), D, default_table(), bsgs_bits(), g_bits, g_frac, nullptr, f.data()) != 0) abort();      // unwrap() panics when there is no log in range
    return leak(f);
}
// ---- :228-258 create_rangeproof -> PyRes{ *PyVec[ proofs, commitments ] } ----------------------------------------------------------------------------
PyRes create_rangeproof(const float *value_ptr, size_t value_len, const uint8_t *blinding_ptr, size_t blinding_len, size_t range_exp, size_t n_partition) {
    need(value_ptr); need(blinding_ptr);
    bytes bl; if (!de_vec32(blinding_ptr, blinding_len, bl)) abort();
    if (bl.size() / 32 != value_len) return err_res("Wrong number of blinding factors supplied.");
    size_t np = 0, plen = 0; rofl_range_proof_shape(value_len, (int)range_exp, n_partition, &np, &plen);
    bytes proofs(np * plen), commits(32 * value_len); uint8_t seed[32]; fresh_seed(seed);
    const int rc = rofl_range_prove(ctx(), value_ptr, bl.data(), value_len, (int)range_exp, n_partition, g_bits, g_frac, seed, proofs.data(), &plen, &np, commits.data());
    if (rc != 0) return err_res(err_text(rc));
    return ok_pyvecs({leak(ser_vec_bytes(proofs.data(), np, plen)), leak(ser_vec32(commits.data(), value_len))});
}
// ---- :265-287 verify_rangeproof(commitments, proofs, range) -> PyRes{ *bool } ------------------------------------------------------------------------
PyRes verify_rangeproof(const uint8_t *commit_ptr, size_t commit_len, const uint8_t *proof_ptr, size_t proof_len, size_t range_exp) {
    need(commit_ptr); need(proof_ptr);
    bytes c, p; size_t plen = 0, np = 0;
    if (!de_vec32(commit_ptr, commit_len, c) || !de_range_proofs(proof_ptr, proof_len, p, plen, np)) abort();
    uint8_t seed[32]; fresh_seed(seed);
    const int rc = rofl_range_verify(ctx(), p.data(), plen, np, c.data(), c.size() / 32, (int)range_exp, seed);
    if (rc < 0) return err_res(err_text(rc));
    return ok_bool(rc == 1);
}
// ---- :295-321 create_randproof -> PyRes{ *PyVec[ Vec<RandProof>, Vec<ElGamalPair> ] } ----------------------------------------------------------------
PyRes create_randproof(const float *value_ptr, size_t value_len, const uint8_t *blinding_ptr, size_t blinding_len) {
    need(value_ptr); need(blinding_ptr);
    bytes bl; if (!de_vec32(blinding_ptr, blinding_len, bl) || bl.size() / 32 != value_len) abort();
    bytes proofs(128 * value_len), pairs(64 * value_len); uint8_t seed[32]; fresh_seed(seed);
    const int rc = rofl_rand_prove(ctx(), value_ptr, nullptr, bl.data(), value_len, g_bits, g_frac, seed, proofs.data(), pairs.data());
    if (rc != 0) return err_res(err_text(rc));
    return ok_pyvecs({leak(ser_vec_bytes(proofs.data(), value_len, 128)), leak(ser_vec_bytes(pairs.data(), value_len, 64))});
}
// ---- :324-370 verify_randproof(pedersen halves, R halves, proofs) (without the reference's debug prints) ----------------------------------------------
PyRes verify_randproof(const uint8_t *ped_commit_ptr, size_t ped_commit_len, const uint8_t *rand_commit_ptr, size_t rand_commit_len, const uint8_t *randproof_ptr, size_t proof_len) {
    need(ped_commit_ptr); need(rand_commit_ptr); need(randproof_ptr);
    bytes L, R, pf; if (!de_vec32(ped_commit_ptr, ped_commit_len, L) || !de_vec32(rand_commit_ptr, rand_commit_len, R) || !de_vec_bytes(randproof_ptr, proof_len, 128, pf)) abort();
    const size_t D = std::min(L.size(), R.size()) / 32;                 // zip()
    if (pf.size() / 128 != D) return err_res("Proof data could not be parsed.");
    bytes pairs(64 * D); for (size_t i = 0; i < D; i++) { memcpy(&pairs[64 * i], &L[32 * i], 32); memcpy(&pairs[64 * i + 32], &R[32 * i], 32); }
    const int rc = rofl_rand_verify(ctx(), pf.data(), pairs.data(), D);
    if (rc < 0) return err_res(err_text(rc));
    return ok_bool(rc == 1);
}
// ---- :373-412 create_squarerandproof -> PyRes{ *PyVec[ Vec<SquareRandProof>, Vec<SquareRandProofCommitments> ] } --------------------------------------
PyRes create_squarerandproof(const float *value_ptr, size_t value_len, const uint8_t *blinding_1_ptr, size_t blinding_1_len, const uint8_t *blinding_2_ptr, size_t blinding_2_len) {
    need(value_ptr); need(blinding_1_ptr); need(blinding_2_ptr);
    bytes b1, b2; if (!de_vec32(blinding_1_ptr, blinding_1_len, b1) || !de_vec32(blinding_2_ptr, blinding_2_len, b2) || b1.size() / 32 != value_len || b2.size() / 32 != value_len) abort();
    bytes proofs(192 * value_len), com(96 * value_len); uint8_t seed[32]; fresh_seed(seed);
    const int rc = rofl_square_rand_prove(ctx(), value_ptr, nullptr, b1.data(), b2.data(), value_len, g_bits, g_frac, seed, proofs.data(), com.data());
    if (rc != 0) return err_res(err_text(rc));
    return ok_pyvecs({leak(ser_vec_bytes(proofs.data(), value_len, 192)), leak(ser_vec_bytes(com.data(), value_len, 96))});
}
// ---- :415-438 verify_squarerandproof(commitments, proofs) ---------------------------------------------------------------------------------------------
PyRes verify_squarerandproof(const uint8_t *commit_ptr, size_t commit_len, const uint8_t *randproof_ptr, size_t proof_len) {
    need(commit_ptr); need(randproof_ptr);
    bytes com, pf; if (!de_vec_bytes(commit_ptr, commit_len, 96, com) || !de_vec_bytes(randproof_ptr, proof_len, 192, pf)) abort();
    if (com.size() / 96 != pf.size() / 192) return err_res("Proof data could not be parsed.");
    const int rc = rofl_square_rand_verify(ctx(), pf.data(), com.data(), com.size() / 96);
    if (rc < 0) return err_res(err_text(rc));
    return ok_bool(rc == 1);
}
// ---- :441-504 create_l2proof -> PyRes{ *PyVec[ Vec<SquareRandProof>, Vec<SquareRandProofCommitments>, RangeProof, RistrettoPoint ] } ---------------------
PyRes create_l2proof(const float *value_ptr, size_t value_len, const uint8_t *blinding_1_ptr, size_t blinding_1_len, const uint8_t *blinding_2_ptr, size_t blinding_2_len, size_t range_exp, size_t n_partition) {
    (void)n_partition;                                                 // create_rangeproof_l2 ignores it too (one proof on the sum, l2_range_proof_vec/mod.rs:15-140)
    need(value_ptr); need(blinding_1_ptr); need(blinding_2_ptr);
    bytes b1, b2; if (!de_vec32(blinding_1_ptr, blinding_1_len, b1) || !de_vec32(blinding_2_ptr, blinding_2_len, b2) || b1.size() / 32 != value_len || b2.size() / 32 != value_len) abort();
    uint8_t seed[32]; fresh_seed(seed);
    bytes rp(rofl_range_proof_len(range_exp ? range_exp : 1)), sq(32); size_t rplen = 0;
    const int rc_range = rofl_l2_prove(ctx(), value_ptr, b2.data(), value_len, (int)range_exp, g_bits, g_frac, seed, rp.data(), &rplen, sq.data());
    bytes proofs(192 * value_len), com(96 * value_len);
    const int rc_rand = rofl_square_rand_prove(ctx(), value_ptr, nullptr, b1.data(), b2.data(), value_len, g_bits, g_frac, seed, proofs.data(), com.data());
    if (rc_rand != 0) return err_res(err_text(rc_rand));
    if (rc_range != 0) return err_res(err_text(rc_range));
    bytes rp_ser; wr64(rp_ser, rplen); rp_ser.insert(rp_ser.end(), rp.begin(), rp.begin() + rplen);
    return ok_pyvecs({leak(ser_vec_bytes(proofs.data(), value_len, 192)), leak(ser_vec_bytes(com.data(), value_len, 96)), leak(rp_ser), leak(sq)});
}
// ---- :507-552 verify_l2proof.  The reference reads a fixed 616 bytes of range proof (8 + 608: a 32-bit sum proof) and 40 bytes of point; here the
//      proof's own bincode length prefix is honoured, so every bit size works, and 32 bytes of point are read --------------------------------------------
PyRes verify_l2proof(const uint8_t *commit_ptr, size_t commit_len, const uint8_t *randproof_ptr, size_t proof_len, const uint8_t *rangeproof_ptr, const uint8_t *square_ptr, size_t prove_range) {
    need(commit_ptr); need(randproof_ptr); need(rangeproof_ptr); need(square_ptr);
    bytes com, pf; if (!de_vec_bytes(commit_ptr, commit_len, 96, com) || !de_vec_bytes(randproof_ptr, proof_len, 192, pf)) abort();
    const size_t D = com.size() / 96; if (D == 0) abort();             // reduce().unwrap()
    const size_t rplen = rd64(rangeproof_ptr);
    bytes csq(32 * D), sum(32); for (size_t i = 0; i < D; i++) memcpy(&csq[32 * i], &com[96 * i + 64], 32);
    if (rofl_aggregate(ctx(), csq.data(), D, 1, 0, sum.data()) != 0) abort();
    if (memcmp(sum.data(), square_ptr, 32) != 0) return err_res("Commitments do not sum to square commitment");      // L2RangeProofError::SumError
    if (pf.size() / 192 != D) return err_res("Proof data could not be parsed.");
    uint8_t seed[32]; fresh_seed(seed);
    const int v1 = rofl_square_rand_verify(ctx(), pf.data(), com.data(), D);
    const int v2 = rofl_l2_verify(ctx(), rangeproof_ptr + 8, rplen, square_ptr, (int)prove_range, seed);
    if (v1 < 0) return err_res(err_text(v1));
    if (v2 < 0) return err_res(err_text(v2));
    return ok_bool(v1 == 1 && v2 == 1);
}
// ---- :555-649 split / join helpers -------------------------------------------------------------------------------------------------------------------
PyVec split_elgamal_pair_vector(const uint8_t *commit_ptr, size_t commit_len) {
    need(commit_ptr); bytes pr; if (!de_vec_bytes(commit_ptr, commit_len, 64, pr)) abort();
    const size_t D = pr.size() / 64; bytes L(32 * D), R(32 * D);
    for (size_t i = 0; i < D; i++) { memcpy(&L[32 * i], &pr[64 * i], 32); memcpy(&R[32 * i], &pr[64 * i + 32], 32); }
    return leak_pyvecs({leak(ser_vec32(L.data(), D)), leak(ser_vec32(R.data(), D))});
}
PyVec join_to_elgamal_pair_vector(const uint8_t *ped_commit_ptr, size_t ped_commit_len, const uint8_t *rand_commit_ptr, size_t rand_commit_len) {
    need(ped_commit_ptr); need(rand_commit_ptr);
    bytes L, R; if (!de_vec32(ped_commit_ptr, ped_commit_len, L) || !de_vec32(rand_commit_ptr, rand_commit_len, R)) abort();
    const size_t D = std::min(L.size(), R.size()) / 32; bytes pr(64 * D);
    for (size_t i = 0; i < D; i++) { memcpy(&pr[64 * i], &L[32 * i], 32); memcpy(&pr[64 * i + 32], &R[32 * i], 32); }
    return leak(ser_vec_bytes(pr.data(), D, 64));
}
PyVec split_squaretriple_pair_vector(const uint8_t *commit_ptr, size_t commit_len) {
    need(commit_ptr); bytes tr; if (!de_vec_bytes(commit_ptr, commit_len, 96, tr)) abort();
    const size_t D = tr.size() / 96; bytes a(32 * D), b(32 * D), c(32 * D);
    for (size_t i = 0; i < D; i++) { memcpy(&a[32 * i], &tr[96 * i], 32); memcpy(&b[32 * i], &tr[96 * i + 32], 32); memcpy(&c[32 * i], &tr[96 * i + 64], 32); }
    return leak_pyvecs({leak(ser_vec32(a.data(), D)), leak(ser_vec32(b.data(), D)), leak(ser_vec32(c.data(), D))});
}
PyVec join_to_squaretriple_pair_vector(const uint8_t *ped_commit_ptr, size_t ped_commit_len, const uint8_t *rand_commit_ptr, size_t rand_commit_len, const uint8_t *square_commit_ptr, size_t square_commit_len) {
    need(ped_commit_ptr); need(rand_commit_ptr);
    bytes a, b, c; if (!de_vec32(ped_commit_ptr, ped_commit_len, a) || !de_vec32(rand_commit_ptr, rand_commit_len, b) || !de_vec32(square_commit_ptr, square_commit_len, c)) abort();
    const size_t D = std::min(std::min(a.size(), b.size()), c.size()) / 32; bytes tr(96 * D);
    for (size_t i = 0; i < D; i++) { memcpy(&tr[96 * i], &a[32 * i], 32); memcpy(&tr[96 * i + 32], &b[32 * i], 32); memcpy(&tr[96 * i + 64], &c[32 * i], 32); }
    return leak(ser_vec_bytes(tr.data(), D, 96));
}
// ---- :652-673 clipping ---------------------------------------------------------------------------------------------------------------------------------
PyVec clip_to_range(const float *value_ptr, size_t value_len, size_t range) {
    need(value_ptr); cfg_from_env(); std::vector<float> o(value_len);
    rofl_clip_f32_to_range_vec(value_ptr, value_len, (int)range, g_bits, g_frac, o.data());
    return leak(o);
}
PyVec quantize_probabilistic(const float *value_ptr, size_t value_len, size_t range) { return clip_to_range(value_ptr, value_len, range); }      // (the reference's is the same clip)
// ---- :675-704 comparisons.  RistrettoPoint equality == equality of the canonical encodings -----------------------------------------------------------------
PyRes commits_equal(const uint8_t *commit_a_ptr, const uint8_t *commit_b_ptr, size_t commit_len) {
    need(commit_a_ptr); need(commit_b_ptr);
    bytes a, b; if (!de_vec32(commit_a_ptr, commit_len, a) || !de_vec32(commit_b_ptr, commit_len, b)) abort();
    return ok_bool(a == b);
}
PyRes equals_neutral_group_element_vec(const uint8_t *commit_ptr, size_t commit_len) {
    need(commit_ptr); bytes a; if (!de_vec32(commit_ptr, commit_len, a)) abort();
    bool z = true; for (uint8_t x : a) z = z && x == 0;                  // the identity encodes as 32 zero bytes
    return ok_bool(z);
}
// ---- :706-724 constant vectors ---------------------------------------------------------------------------------------------------------------------------
PyVec create_zero_scalar_vector(size_t len) { bytes z(32 * len, 0); return leak(ser_vec32(z.data(), len)); }
PyVec create_zero_group_element_vector(size_t len) { bytes z(32 * len, 0); return leak(ser_vec32(z.data(), len)); }
PyVec create_random_blinding_vector(size_t len) { bytes v(32 * len); uint8_t seed[32]; fresh_seed(seed); rofl_rnd_scalar_vec(seed, len, v.data()); return leak(ser_vec32(v.data(), len)); }
// ---- :727-735 add_scalars: the sum of a scalar vector as one bincode Scalar (32 bytes) -------------------------------------------------------------------------
PyVec add_scalars(const uint8_t *commit_ptr, size_t commit_len) {
    need(commit_ptr); bytes v; if (!de_vec32(commit_ptr, commit_len, v) || v.empty()) abort();
    bytes s(v.begin(), v.begin() + 32);
    for (size_t i = 1; i < v.size() / 32; i++) rofl_scalar_ops(0, s.data(), &v[32 * i], 1, s.data());
    return leak(s);
}
// ---- :737-764 filter_unequal_commits -> PyVec[ left, right ] of the positions where a != b ----------------------------------------------------------------------
PyVec filter_unequal_commits(const uint8_t *commit_a_ptr, const uint8_t *commit_b_ptr, size_t commit_len) {
    need(commit_a_ptr); need(commit_b_ptr);
    bytes a, b; if (!de_vec32(commit_a_ptr, commit_len, a) || !de_vec32(commit_b_ptr, commit_len, b)) abort();
    bytes l, r; const size_t D = std::min(a.size(), b.size()) / 32;
    for (size_t i = 0; i < D; i++) if (memcmp(&a[32 * i], &b[32 * i], 32) != 0) { l.insert(l.end(), &a[32 * i], &a[32 * i] + 32); r.insert(r.end(), &b[32 * i], &b[32 * i] + 32); }
    return leak_pyvecs({leak(ser_vec32(l.data(), l.size() / 32)), leak(ser_vec32(r.data(), r.size() / 32))});
}
}  // extern "C"
