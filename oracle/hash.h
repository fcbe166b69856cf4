/* ORACLE -- TEST INFRASTRUCTURE ONLY (see fe51.h).
 *
 * Keccak-f[1600], SHA3-512, SHA3-256, SHAKE256 (FIPS 202), STROBE-128 as used by Merlin 3.0.0,
 * the Merlin transcript, and the ChaCha20 keystream of rand_chacha 0.3 `ChaCha20Rng`
 * (all third-party, Cargo.lock:761-762,830-831,1141-1164,1419-1420).
 * Reference call sites: `Transcript::new(b"RangeProof")` range_proof_vec/mod.rs:124,200;
 * `b"L2RangeProof"` l2_range_proof_vec/mod.rs:163,238; `b"SquareProof"` square_proof_vec/mod.rs:52,106;
 * `hash_from_bytes::<Sha3_512>` el_gamal.rs:35-37; extension trait rand_proof/transcript.rs:10-45.
 * Pinned by hashlib (SHA3/SHAKE) and the published Merlin conformance vector in tests/.
 */
#ifndef ROFL_ORACLE_HASH_H
#define ROFL_ORACLE_HASH_H
#include <stdint.h>
#include <string.h>
#include <stddef.h>

static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static inline uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
static inline void keccak_f1600(uint64_t st[25]) {
    uint64_t bc[5], t;
    for (int r = 0; r < 24; r++) {
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) { t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1); for (int j = 0; j < 25; j += 5) st[j + i] ^= t; }
        t = st[1];
        for (int i = 0; i < 24; i++) { int j = KECCAK_PIL[i]; bc[0] = st[j]; st[j] = rotl64(t, KECCAK_ROT[i]); t = bc[0]; }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KECCAK_RC[r];
    }
}

/* generic sponge (little-endian host assumed: state bytes alias the u64 lanes) */
typedef struct { uint64_t st[25]; int rate, pos; } sponge;
static inline void sponge_init(sponge *s, int rate) { memset(s, 0, sizeof *s); s->rate = rate; }
static inline void sponge_absorb(sponge *s, const uint8_t *in, size_t n) {
    uint8_t *b = (uint8_t *)s->st;
    for (size_t i = 0; i < n; i++) { b[s->pos++] ^= in[i]; if (s->pos == s->rate) { keccak_f1600(s->st); s->pos = 0; } }
}
static inline void sponge_finish(sponge *s, uint8_t dsuffix) {
    uint8_t *b = (uint8_t *)s->st;
    b[s->pos] ^= dsuffix; b[s->rate - 1] ^= 0x80; keccak_f1600(s->st); s->pos = 0;
}
static inline void sponge_squeeze(sponge *s, uint8_t *out, size_t n) {
    uint8_t *b = (uint8_t *)s->st;
    for (size_t i = 0; i < n; i++) { if (s->pos == s->rate) { keccak_f1600(s->st); s->pos = 0; } out[i] = b[s->pos++]; }
}
static inline void sha3_512(uint8_t out[64], const uint8_t *in, size_t n) {
    sponge s; sponge_init(&s, 72); sponge_absorb(&s, in, n); sponge_finish(&s, 0x06); sponge_squeeze(&s, out, 64);
}
static inline void sha3_256(uint8_t out[32], const uint8_t *in, size_t n) {
    sponge s; sponge_init(&s, 136); sponge_absorb(&s, in, n); sponge_finish(&s, 0x06); sponge_squeeze(&s, out, 32);
}
/* SHAKE256: init(136), absorb..., finish(0x1f), squeeze... */

/* ---- STROBE-128 / Merlin ---------------------------------------------------- */
#define STROBE_R 166
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
typedef struct { uint64_t st[25]; uint8_t pos, pos_begin, cur_flags; } strobe;

static inline void strobe_run_f(strobe *s) {
    uint8_t *b = (uint8_t *)s->st;
    b[s->pos] ^= s->pos_begin; b[s->pos + 1] ^= 0x04; b[STROBE_R + 1] ^= 0x80;
    keccak_f1600(s->st); s->pos = 0; s->pos_begin = 0;
}
static inline void strobe_absorb(strobe *s, const uint8_t *d, size_t n) {
    uint8_t *b = (uint8_t *)s->st;
    for (size_t i = 0; i < n; i++) { b[s->pos++] ^= d[i]; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static inline void strobe_squeeze(strobe *s, uint8_t *d, size_t n) {
    uint8_t *b = (uint8_t *)s->st;
    for (size_t i = 0; i < n; i++) { d[i] = b[s->pos]; b[s->pos] = 0; s->pos++; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static inline void strobe_begin_op(strobe *s, uint8_t flags, int more) {
    if (more) return;                                  /* continuation of the same operation */
    uint8_t old_begin = s->pos_begin;
    s->pos_begin = s->pos + 1; s->cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    strobe_absorb(s, hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && s->pos != 0) strobe_run_f(s);
}
static inline void strobe_meta_ad(strobe *s, const uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_M | FLAG_A, more); strobe_absorb(s, d, n); }
static inline void strobe_ad(strobe *s, const uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_A, more); strobe_absorb(s, d, n); }
static inline void strobe_prf(strobe *s, uint8_t *d, size_t n, int more) { strobe_begin_op(s, FLAG_I | FLAG_A | FLAG_C, more); strobe_squeeze(s, d, n); }
static inline void strobe_init(strobe *s, const uint8_t *label, size_t n) {
    memset(s, 0, sizeof *s);
    uint8_t *b = (uint8_t *)s->st;
    static const uint8_t hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(b, hdr, 6); memcpy(b + 6, "STROBEv1.0.2", 12);
    keccak_f1600(s->st);
    strobe_meta_ad(s, label, n, 0);
}

typedef strobe transcript;
static inline void u32le(uint8_t b[4], uint32_t x) { b[0] = x; b[1] = x >> 8; b[2] = x >> 16; b[3] = x >> 24; }
static inline void transcript_append(transcript *t, const char *label, const uint8_t *msg, size_t n) {
    uint8_t len[4]; u32le(len, (uint32_t)n);
    strobe_meta_ad(t, (const uint8_t *)label, strlen(label), 0);
    strobe_meta_ad(t, len, 4, 1);
    strobe_ad(t, msg, n, 0);
}
/* labels that are raw bytes (the compressed-rand-proof per-index 3-byte labels) */
static inline void transcript_append_l(transcript *t, const uint8_t *label, size_t ll, const uint8_t *msg, size_t n) {
    uint8_t len[4]; u32le(len, (uint32_t)n);
    strobe_meta_ad(t, label, ll, 0); strobe_meta_ad(t, len, 4, 1); strobe_ad(t, msg, n, 0);
}
static inline void transcript_init(transcript *t, const char *label) {
    strobe_init(t, (const uint8_t *)"Merlin v1.0", 11);
    transcript_append(t, "dom-sep", (const uint8_t *)label, strlen(label));
}
static inline void transcript_append_u64(transcript *t, const char *label, uint64_t x) {
    uint8_t b[8]; for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    transcript_append(t, label, b, 8);
}
static inline void transcript_challenge(transcript *t, const char *label, uint8_t *out, size_t n) {
    uint8_t len[4]; u32le(len, (uint32_t)n);
    strobe_meta_ad(t, (const uint8_t *)label, strlen(label), 0);
    strobe_meta_ad(t, len, 4, 1);
    strobe_prf(t, out, n, 0);
}

/* ---- ChaCha20 keystream, 64-bit block counter, zero nonce (rand_chacha ChaCha20Rng) -------- */
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define CHACHA_QR(a, b, c, d) \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);
static inline void chacha20_block(uint8_t out[64], const uint8_t key[32], uint64_t counter) {
    uint32_t in[16], x[16];
    in[0] = 0x61707865; in[1] = 0x3320646e; in[2] = 0x79622d32; in[3] = 0x6b206574;
    memcpy(in + 4, key, 32);
    in[12] = (uint32_t)counter; in[13] = (uint32_t)(counter >> 32); in[14] = 0; in[15] = 0;
    memcpy(x, in, 64);
    for (int i = 0; i < 10; i++) {
        CHACHA_QR(x[0], x[4], x[8], x[12]) CHACHA_QR(x[1], x[5], x[9], x[13])
        CHACHA_QR(x[2], x[6], x[10], x[14]) CHACHA_QR(x[3], x[7], x[11], x[15])
        CHACHA_QR(x[0], x[5], x[10], x[15]) CHACHA_QR(x[1], x[6], x[11], x[12])
        CHACHA_QR(x[2], x[7], x[8], x[13]) CHACHA_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) x[i] += in[i];
    memcpy(out, x, 64);
}
#endif
