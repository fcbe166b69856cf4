// Keccak-f[1600] on 64-bit lanes, SHA3-256/512, SHAKE256, STROBE-128 + Merlin 3.0 transcripts and the
// ChaCha20 keystream (rand_chacha ChaCha20Rng layout), all __host__ __device__:
//   * device: per-element Sigma-proof transcripts (square_proof_vec/mod.rs:46-59 runs one
//     `Transcript::new(b"SquareProof")` per element), generator chains (bulletproofs GeneratorsChain),
//     nonce streams;
//   * host: the sequential Fiat-Shamir transcript of each Bulletproof chunk (range_proof_vec/mod.rs:124,200).
// State is kept as 25 u64 lanes; byte access goes through shifts so the same code runs on either side.
#pragma once
#include "fe25519.cuh"
#include <string.h>

#if defined(__CUDA_ARCH__)
#define HASH_CONST static __device__ __constant__ const
#else
#define HASH_CONST static const
#endif
HASH_CONST uint64_t KECCAK_RC_[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

HD uint64_t rotl64_(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// fully unrolled round so every lane lives in a register on the device
HDNI void keccak_f1600(uint64_t s[25]) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 0; r < 24; r++) {
        uint64_t c0 = s[0] ^ s[5] ^ s[10] ^ s[15] ^ s[20], c1 = s[1] ^ s[6] ^ s[11] ^ s[16] ^ s[21];
        uint64_t c2 = s[2] ^ s[7] ^ s[12] ^ s[17] ^ s[22], c3 = s[3] ^ s[8] ^ s[13] ^ s[18] ^ s[23];
        uint64_t c4 = s[4] ^ s[9] ^ s[14] ^ s[19] ^ s[24];
        uint64_t d0 = c4 ^ rotl64_(c1, 1), d1 = c0 ^ rotl64_(c2, 1), d2 = c1 ^ rotl64_(c3, 1), d3 = c2 ^ rotl64_(c4, 1), d4 = c3 ^ rotl64_(c0, 1);
        // theta + rho + pi
        uint64_t b0 = s[0] ^ d0;
        uint64_t b1 = rotl64_(s[6] ^ d1, 44), b2 = rotl64_(s[12] ^ d2, 43), b3 = rotl64_(s[18] ^ d3, 21), b4 = rotl64_(s[24] ^ d4, 14);
        uint64_t b5 = rotl64_(s[3] ^ d3, 28), b6 = rotl64_(s[9] ^ d4, 20), b7 = rotl64_(s[10] ^ d0, 3), b8 = rotl64_(s[16] ^ d1, 45), b9 = rotl64_(s[22] ^ d2, 61);
        uint64_t b10 = rotl64_(s[1] ^ d1, 1), b11 = rotl64_(s[7] ^ d2, 6), b12 = rotl64_(s[13] ^ d3, 25), b13 = rotl64_(s[19] ^ d4, 8), b14 = rotl64_(s[20] ^ d0, 18);
        uint64_t b15 = rotl64_(s[4] ^ d4, 27), b16 = rotl64_(s[5] ^ d0, 36), b17 = rotl64_(s[11] ^ d1, 10), b18 = rotl64_(s[17] ^ d2, 15), b19 = rotl64_(s[23] ^ d3, 56);
        uint64_t b20 = rotl64_(s[2] ^ d2, 62), b21 = rotl64_(s[8] ^ d3, 55), b22 = rotl64_(s[14] ^ d4, 39), b23 = rotl64_(s[15] ^ d0, 41), b24 = rotl64_(s[21] ^ d1, 2);
        // chi + iota
        s[0] = b0 ^ (~b1 & b2) ^ KECCAK_RC_[r]; s[1] = b1 ^ (~b2 & b3); s[2] = b2 ^ (~b3 & b4); s[3] = b3 ^ (~b4 & b0); s[4] = b4 ^ (~b0 & b1);
        s[5] = b5 ^ (~b6 & b7); s[6] = b6 ^ (~b7 & b8); s[7] = b7 ^ (~b8 & b9); s[8] = b8 ^ (~b9 & b5); s[9] = b9 ^ (~b5 & b6);
        s[10] = b10 ^ (~b11 & b12); s[11] = b11 ^ (~b12 & b13); s[12] = b12 ^ (~b13 & b14); s[13] = b13 ^ (~b14 & b10); s[14] = b14 ^ (~b10 & b11);
        s[15] = b15 ^ (~b16 & b17); s[16] = b16 ^ (~b17 & b18); s[17] = b17 ^ (~b18 & b19); s[18] = b18 ^ (~b19 & b15); s[19] = b19 ^ (~b15 & b16);
        s[20] = b20 ^ (~b21 & b22); s[21] = b21 ^ (~b22 & b23); s[22] = b22 ^ (~b23 & b24); s[23] = b23 ^ (~b24 & b20); s[24] = b24 ^ (~b20 & b21);
    }
}

HD void st_xor_byte(uint64_t st[25], int pos, uint8_t b) { st[pos >> 3] ^= (uint64_t)b << (8 * (pos & 7)); }
HD uint8_t st_get_byte(const uint64_t st[25], int pos) { return (uint8_t)(st[pos >> 3] >> (8 * (pos & 7))); }
HD void st_clear_byte(uint64_t st[25], int pos) { st[pos >> 3] &= ~(0xffULL << (8 * (pos & 7))); }

// ---- generic sponge -------------------------------------------------------------------------------
struct sponge { uint64_t st[25]; int rate, pos; };
HD void sponge_init(sponge &s, int rate) { for (int i = 0; i < 25; i++) s.st[i] = 0; s.rate = rate; s.pos = 0; }
HD void sponge_absorb(sponge &s, const uint8_t *in, size_t n) {
    for (size_t i = 0; i < n; i++) { st_xor_byte(s.st, s.pos++, in[i]); if (s.pos == s.rate) { keccak_f1600(s.st); s.pos = 0; } }
}
HD void sponge_finish(sponge &s, uint8_t dsuffix) { st_xor_byte(s.st, s.pos, dsuffix); st_xor_byte(s.st, s.rate - 1, 0x80); keccak_f1600(s.st); s.pos = 0; }
HD void sponge_squeeze(sponge &s, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) { if (s.pos == s.rate) { keccak_f1600(s.st); s.pos = 0; } out[i] = st_get_byte(s.st, s.pos++); }
}
HD void sha3_512(uint8_t out[64], const uint8_t *in, size_t n) { sponge s; sponge_init(s, 72); sponge_absorb(s, in, n); sponge_finish(s, 0x06); sponge_squeeze(s, out, 64); }
HD void sha3_256(uint8_t out[32], const uint8_t *in, size_t n) { sponge s; sponge_init(s, 136); sponge_absorb(s, in, n); sponge_finish(s, 0x06); sponge_squeeze(s, out, 32); }

// ---- STROBE-128 as used by Merlin 3.0 ----------------------------------------------
#define STROBE_R 166
struct strobe { uint64_t st[25]; uint8_t pos, pos_begin, cur_flags; };
HD void strobe_run_f(strobe &s) {
    st_xor_byte(s.st, s.pos, s.pos_begin); st_xor_byte(s.st, s.pos + 1, 0x04); st_xor_byte(s.st, STROBE_R + 1, 0x80);
    keccak_f1600(s.st); s.pos = 0; s.pos_begin = 0;
}
HD void strobe_absorb(strobe &s, const uint8_t *d, size_t n) {
    size_t i = 0;
#if !defined(__CUDA_ARCH__) && defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
    // host: whole 64-bit lanes at a time once the position is lane aligned (a verifier absorbs 1024 commitments per chunk)
    while (i < n && (s.pos & 7)) { st_xor_byte(s.st, s.pos++, d[i++]); if (s.pos == STROBE_R) strobe_run_f(s); }
    while (n - i >= 8) {
        if (s.pos + 8 > STROBE_R) {                       // the rate (166) ends inside a lane: finish the block byte by byte
            while (s.pos != 0 && i < n) { st_xor_byte(s.st, s.pos++, d[i++]); if (s.pos == STROBE_R) strobe_run_f(s); }
            continue;
        }
        uint64_t w; memcpy(&w, d + i, 8); s.st[s.pos >> 3] ^= w; s.pos += 8; i += 8;
    }
#endif
    for (; i < n; i++) { st_xor_byte(s.st, s.pos++, d[i]); if (s.pos == STROBE_R) strobe_run_f(s); }
}
HD void strobe_squeeze(strobe &s, uint8_t *d, size_t n) {
    for (size_t i = 0; i < n; i++) { d[i] = st_get_byte(s.st, s.pos); st_clear_byte(s.st, s.pos); s.pos++; if (s.pos == STROBE_R) strobe_run_f(s); }
}
HD void strobe_begin_op(strobe &s, uint8_t flags) {
    uint8_t hdr[2] = {s.pos_begin, flags};
    s.pos_begin = s.pos + 1; s.cur_flags = flags;
    strobe_absorb(s, hdr, 2);
    if ((flags & (4 | 32)) && s.pos != 0) strobe_run_f(s);          // C or K
}
HD void strobe_meta_ad(strobe &s, const uint8_t *d, size_t n, bool more) { if (!more) strobe_begin_op(s, 16 | 2); strobe_absorb(s, d, n); }
HD void strobe_ad(strobe &s, const uint8_t *d, size_t n, bool more) { if (!more) strobe_begin_op(s, 2); strobe_absorb(s, d, n); }
HD void strobe_prf(strobe &s, uint8_t *d, size_t n, bool more) { if (!more) strobe_begin_op(s, 1 | 2 | 4); strobe_squeeze(s, d, n); }
HD void strobe_init(strobe &s, const uint8_t *label, size_t n) {
    for (int i = 0; i < 25; i++) s.st[i] = 0;
    const uint8_t hdr[18] = {1, STROBE_R + 2, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
    for (int i = 0; i < 18; i++) st_xor_byte(s.st, i, hdr[i]);
    keccak_f1600(s.st);
    s.pos = 0; s.pos_begin = 0; s.cur_flags = 0;
    strobe_meta_ad(s, label, n, false);
}
typedef strobe transcript;
HD size_t cstrlen_(const char *s) { size_t n = 0; while (s[n]) n++; return n; }
HD void transcript_append(transcript &t, const char *label, const uint8_t *msg, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(t, (const uint8_t *)label, cstrlen_(label), false);
    strobe_meta_ad(t, len, 4, true);
    strobe_ad(t, msg, n, false);
}
// label given as raw bytes (the compressed rand proof labels its pairs with 3 arbitrary bytes, zeros included)
HD void transcript_append_l(transcript &t, const uint8_t *label, size_t ll, const uint8_t *msg, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(t, label, ll, false);
    strobe_meta_ad(t, len, 4, true);
    strobe_ad(t, msg, n, false);
}
HD void transcript_init(transcript &t, const char *label) {
    const uint8_t proto[11] = {'M', 'e', 'r', 'l', 'i', 'n', ' ', 'v', '1', '.', '0'};
    strobe_init(t, proto, 11);
    transcript_append(t, "dom-sep", (const uint8_t *)label, cstrlen_(label));
}
HD void transcript_append_u64(transcript &t, const char *label, uint64_t x) {
    uint8_t b[8]; for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    transcript_append(t, label, b, 8);
}
HD void transcript_challenge(transcript &t, const char *label, uint8_t *out, size_t n) {
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(t, (const uint8_t *)label, cstrlen_(label), false);
    strobe_meta_ad(t, len, 4, true);
    strobe_prf(t, out, n, false);
}

// ---- ChaCha20 block, 64-bit counter, zero nonce; output as 16 little-endian words --------------------
HD uint32_t rotl32_(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define CHACHA_QR_(a, b, c, d) \
    a += b; d ^= a; d = rotl32_(d, 16); c += d; b ^= c; b = rotl32_(b, 12); \
    a += b; d ^= a; d = rotl32_(d, 8);  c += d; b ^= c; b = rotl32_(b, 7);
HD void chacha20_block_words(uint32_t out[16], const uint32_t key[8], uint64_t counter) {
    uint32_t in[16], x[16];
    in[0] = 0x61707865; in[1] = 0x3320646e; in[2] = 0x79622d32; in[3] = 0x6b206574;
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = (uint32_t)counter; in[13] = (uint32_t)(counter >> 32); in[14] = 0; in[15] = 0;
    for (int i = 0; i < 16; i++) x[i] = in[i];
    for (int i = 0; i < 10; i++) {
        CHACHA_QR_(x[0], x[4], x[8], x[12]) CHACHA_QR_(x[1], x[5], x[9], x[13])
        CHACHA_QR_(x[2], x[6], x[10], x[14]) CHACHA_QR_(x[3], x[7], x[11], x[15])
        CHACHA_QR_(x[0], x[5], x[10], x[15]) CHACHA_QR_(x[1], x[6], x[11], x[12])
        CHACHA_QR_(x[2], x[7], x[8], x[13]) CHACHA_QR_(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}
