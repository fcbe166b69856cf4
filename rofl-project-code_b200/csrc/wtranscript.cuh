// Warp-cooperative Keccak-f / STROBE-128 / Merlin: the transcript machinery shared by the transcript kernels (ts_kernels.cuh) and the fused
// tail kernel of the inner-product argument (kernels.cuh K6b).  Included by kernels.cuh.
#pragma once

#ifdef ROFL_EMUL
#define TS_WSYNC() __syncthreads()
#else
#define TS_WSYNC() __syncwarp()
#endif
#define TS_THREADS 32

// gather form of rho + pi: B[d] = rotl(A[KC_SRC_[d]] ^ D[KC_SRC_[d] % 5], KC_ROT_[d])
HASH_CONST uint8_t KC_SRC_[25] = {0, 6, 12, 18, 24, 3, 9, 10, 16, 22, 1, 7, 13, 19, 20, 4, 5, 11, 17, 23, 2, 8, 14, 15, 21};
HASH_CONST uint8_t KC_ROT_[25] = {0, 44, 43, 21, 14, 28, 20, 3, 45, 61, 1, 6, 25, 8, 18, 27, 36, 10, 15, 56, 62, 55, 39, 41, 2};

// STROBE state of one transcript in shared memory while a warp works on it
#ifdef ROFL_EMUL
struct strobe_sh { uint64_t st[25]; uint64_t C[5]; uint64_t B[25]; uint32_t pos, pos_begin, cur_flags; uint8_t io[64]; };
#else
#define TS_RING 2048                   // bytes of absorb stream composed ahead in shared memory (ring, indexed by absolute stream position)
struct strobe_sh { uint64_t st[25]; uint32_t pos, pos_begin, cur_flags; uint8_t io[64]; uint64_t ring[TS_RING / 8]; };
#endif

#if defined(KG_TS) || defined(KG_FOLD)
#ifdef ROFL_EMUL
// Keccak-f[1600], lane t < 25 owns word t; every lane of the warp must call
DEV void keccak_coop(strobe_sh &h, int lane) {
    const int x = lane % 5, row = lane - x;
    const int src = lane < 25 ? KC_SRC_[lane] : 0, rot = lane < 25 ? KC_ROT_[lane] : 0, sx = src % 5;
    const int d1 = (sx + 4) % 5, d2 = (sx + 1) % 5, c1 = row + (x + 1) % 5, c2 = row + (x + 2) % 5;
    for (int r = 0; r < 24; r++) {
        if (lane < 5) h.C[lane] = h.st[lane] ^ h.st[lane + 5] ^ h.st[lane + 10] ^ h.st[lane + 15] ^ h.st[lane + 20];
        TS_WSYNC();
        if (lane < 25) {
            const uint64_t v = h.st[src] ^ h.C[d1] ^ rotl64_(h.C[d2], 1);
            h.B[lane] = rot ? rotl64_(v, rot) : v;
        }
        TS_WSYNC();
        if (lane < 25) {
            uint64_t v = h.B[lane] ^ (~h.B[c1] & h.B[c2]);
            if (lane == 0) v ^= KECCAK_RC_[r];
            h.st[lane] = v;
        }
        TS_WSYNC();
    }
}
#else
// Keccak-f[1600] over a warp: lane t < 25 holds word t = A[x, y] (t = x + 5 y) in a register; every lane of the warp must call.  Per round 9
// 64-bit shuffles in THREE dependent steps: (1) the four other words of the column (-> column parity c, known to every lane of the column)
// together with the word rho+pi will move here; (2) the parities of the two neighbour columns OF THAT SOURCE word; (3) the two chi operands.
DEV uint64_t keccak_shfl(uint64_t a, int lane) {
    const int x = lane % 5, row = lane - x;
    const int src = lane < 25 ? KC_SRC_[lane] : 0, rot = lane < 25 ? KC_ROT_[lane] : 0, sx = src % 5;
    const int l5 = (lane + 5) % 25, l10 = (lane + 10) % 25, l15 = (lane + 15) % 25, l20 = (lane + 20) % 25;
    const int dm = (sx + 4) % 5, dp = (sx + 1) % 5, c1 = row + (x + 1) % 5, c2 = row + (x + 2) % 5;
    auto sh64 = [](uint64_t v, int from) { const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, from), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), from); return ((uint64_t)hi << 32) | lo; };
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        const uint64_t p1 = sh64(a, l5), p2 = sh64(a, l10), p3 = sh64(a, l15), p4 = sh64(a, l20), as = sh64(a, src);
        const uint64_t c = a ^ p1 ^ p2 ^ p3 ^ p4;                                     // parity of column x (lanes 0..4 hold columns 0..4)
        const uint64_t cm = sh64(c, dm), cp = sh64(c, dp);
        uint64_t b = as ^ cm ^ ((cp << 1) | (cp >> 63));                              // theta applied to the source word
        b = (b << rot) | (b >> ((64 - rot) & 63));                                    // rho (pi is the choice of src)
        const uint64_t b1 = sh64(b, c1), b2 = sh64(b, c2);
        a = b ^ (~b1 & b2);                                                           // chi
        if (lane == 0) a ^= KECCAK_RC_[r];                                            // iota
    }
    return a;
}
#endif
// ---- STROBE / Merlin operations carried out by a whole warp on a state in shared memory -----------------------------------------------
// The 200 state bytes live in strobe_sh::st, the lanes XOR / read message bytes side by side (one byte per lane), the permutation is the warp
// one above (a lone lane running keccak_f1600 on a state in local memory needs 8-15 us per permutation, the warp 2 us).  Control flow is
// uniform: EVERY lane of the warp calls these with the same arguments; `d` must be readable by all lanes (or hold the same bytes in each).
struct wts { uint32_t pos, pos_begin, cur_flags; };                                  // warp-uniform registers
DEV void wt_permute(strobe_sh &h, int lane) {
#ifdef ROFL_EMUL
    keccak_coop(h, lane);
#else
    __syncwarp();
    uint64_t a = lane < 25 ? h.st[lane] : 0;
    a = keccak_shfl(a, lane);
    if (lane < 25) h.st[lane] = a;
    __syncwarp();
#endif
}
DEV void wt_run_f(strobe_sh &h, wts &w, int lane) {
    TS_WSYNC();
    if (lane == 0) { uint8_t *st8 = (uint8_t *)h.st; st8[w.pos] ^= (uint8_t)w.pos_begin; st8[w.pos + 1] ^= 0x04; st8[STROBE_R + 1] ^= 0x80; }
    TS_WSYNC();
    wt_permute(h, lane);
    w.pos = 0; w.pos_begin = 0;
}
DEV void wt_absorb(strobe_sh &h, wts &w, int lane, const uint8_t *d, uint32_t n) {
    uint8_t *st8 = (uint8_t *)h.st;
    for (uint32_t off = 0; off < n;) {
        const uint32_t room = STROBE_R - w.pos, chunk = n - off < room ? n - off : room;
        for (uint32_t i = lane; i < chunk; i += TS_THREADS) st8[w.pos + i] ^= d[off + i];
        w.pos += chunk; off += chunk;
        if (w.pos == STROBE_R) wt_run_f(h, w, lane);
    }
    TS_WSYNC();
}
DEV void wt_squeeze(strobe_sh &h, wts &w, int lane, uint8_t *out, uint32_t n) {       // out: shared memory
    uint8_t *st8 = (uint8_t *)h.st;
    for (uint32_t off = 0; off < n;) {
        const uint32_t room = STROBE_R - w.pos, chunk = n - off < room ? n - off : room;
        for (uint32_t i = lane; i < chunk; i += TS_THREADS) { out[off + i] = st8[w.pos + i]; st8[w.pos + i] = 0; }
        w.pos += chunk; off += chunk;
        if (w.pos == STROBE_R) wt_run_f(h, w, lane);
    }
    TS_WSYNC();
}
DEV void wt_begin_op(strobe_sh &h, wts &w, int lane, uint8_t flags) {
    const uint8_t hdr[2] = {(uint8_t)w.pos_begin, flags};
    w.pos_begin = w.pos + 1; w.cur_flags = flags;
    wt_absorb(h, w, lane, hdr, 2);
    if ((flags & (4 | 32)) && w.pos != 0) wt_run_f(h, w, lane);
}
DEV uint32_t wt_strlen(const char *s) { uint32_t n = 0; while (s[n]) n++; return n; }
DEV void wt_append(strobe_sh &h, wts &w, int lane, const char *label, const uint8_t *msg, uint32_t n) {
    const uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    wt_begin_op(h, w, lane, 16 | 2); wt_absorb(h, w, lane, (const uint8_t *)label, wt_strlen(label)); wt_absorb(h, w, lane, len, 4);
    wt_begin_op(h, w, lane, 2); wt_absorb(h, w, lane, msg, n);
}
DEV void wt_append_u64(strobe_sh &h, wts &w, int lane, const char *label, uint64_t x) {
    uint8_t b[8]; for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    wt_append(h, w, lane, label, b, 8);
}
DEV void wt_challenge(strobe_sh &h, wts &w, int lane, const char *label, uint8_t *out, uint32_t n) {      // out: shared memory
    const uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    wt_begin_op(h, w, lane, 16 | 2); wt_absorb(h, w, lane, (const uint8_t *)label, wt_strlen(label)); wt_absorb(h, w, lane, len, 4);
    wt_begin_op(h, w, lane, 1 | 2 | 4); wt_squeeze(h, w, lane, out, n);
}
// challenge scalar (64 bytes reduced mod l) through h.io; every lane gets it
DEV void wt_challenge_sc(strobe_sh &h, wts &w, int lane, const char *label, sc &out) {
    wt_challenge(h, w, lane, label, h.io, 64);
    uint8_t b[64]; for (int i = 0; i < 64; i++) b[i] = h.io[i];
    TS_WSYNC();
    sc_from_bytes_wide(out, b);
}
// 32 bytes from global memory -> h.io (all lanes see them afterwards); returns whether they are all zero
DEV bool wt_load32(strobe_sh &h, int lane, const uint8_t *g, uint32_t at = 0) {
    TS_WSYNC();
    h.io[at + lane] = g[lane];
    TS_WSYNC();
    uint8_t z = 0; for (int i = 0; i < 32; i++) z |= h.io[at + i];
    return z == 0;
}
DEV void wt_load(strobe_sh &h, wts &w, int lane, const transcript &t) {
    TS_WSYNC();
    if (lane < 25) h.st[lane] = t.st[lane];
    w.pos = t.pos; w.pos_begin = t.pos_begin; w.cur_flags = t.cur_flags;
    TS_WSYNC();
}
DEV void wt_store(const strobe_sh &h, const wts &w, int lane, transcript &t) {
    TS_WSYNC();
    if (lane < 25) t.st[lane] = h.st[lane];
    if (lane == 0) { t.pos = (uint8_t)w.pos; t.pos_begin = (uint8_t)w.pos_begin; t.cur_flags = (uint8_t)w.cur_flags; }
}
DEV void wt_init(strobe_sh &h, wts &w, int lane, const char *label) {
    const uint8_t hdr[18] = {1, STROBE_R + 2, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
    TS_WSYNC();
    if (lane < 25) h.st[lane] = 0;
    TS_WSYNC();
    if (lane < 18) ((uint8_t *)h.st)[lane] ^= hdr[lane];
    TS_WSYNC();
    wt_permute(h, lane);
    w.pos = 0; w.pos_begin = 0; w.cur_flags = 0;
    wt_begin_op(h, w, lane, 16 | 2); wt_absorb(h, w, lane, (const uint8_t *)"Merlin v1.0", 11);
    wt_append(h, w, lane, "dom-sep", (const uint8_t *)label, wt_strlen(label));
}
// m x Transcript::append_message(label (1 byte), 32-byte message) = per message the 41 stream bytes
//   [pos_begin'] [M|A = 0x12] [label] [32 0 0 0]   [pos_begin''] [A = 0x02] [32 message bytes]
// where the two pos_begin bytes are (begin position of the PREVIOUS operation) + 1 if that operation began in the current sponge
// block and 0 otherwise (Strobe128::begin_op / run_f).  Byte k of the stream lands at absolute position A0 + k
// (A0 = position at entry), so every byte is a function of k alone and the lanes fill a 166-byte block together.
// byte at absolute stream position A (which lies in the sponge block starting at `base`)
// (msgs points at message `joff`: the whole array with joff = 0, or the staged window)
DEV uint8_t absorb_stream_byte(uint32_t A, uint32_t base, uint32_t A0, uint8_t pb0, uint8_t label, const uint8_t *msgs, uint32_t joff = 0) {
    const uint32_t k = A - A0, j = k / 41u, r = k - 41u * j, a1 = A0 + 41u * j;
    if (r >= 9) return msgs[32 * (size_t)(j - joff) + (r - 9)];
    if (r == 0) { const uint32_t ap = a1 - 34; return j == 0 ? pb0 : (ap >= base ? (uint8_t)(ap - base + 1) : 0); }      // a1 = A lies in this block
    if (r == 1) return 0x12;
    if (r == 2) return label;
    if (r == 3) return 32;
    if (r < 7) return 0;
    if (r == 7) return a1 >= base ? (uint8_t)(a1 - base + 1) : 0;                                                         // a2 = A lies in this block
    return 0x02;
}
// pos_begin at the run_f that closes the block starting at `base`: (begin of the last operation) + 1 if it lies in this block
DEV uint8_t absorb_close_pb(uint32_t base, uint32_t A0) {
    const uint32_t last = base + (STROBE_R - 1), k = last - A0, j = k / 41u, r = k - 41u * j, a = A0 + 41u * j + (r >= 7 ? 7 : 0);
    return a >= base ? (uint8_t)(a - base + 1) : 0;
}
DEV void strobe_absorb_many_coop(strobe_sh &h, int lane, uint8_t label, const uint8_t *msgs, uint32_t m) {
    if (m == 0) return;
    // 32-bit positions (m < 2^26 commitments per chunk): constant divisions become multiply-shifts
    const uint32_t A0 = h.pos, A_end = A0 + 41u * m, nfull = A_end / STROBE_R;
    const uint8_t pb0 = (uint8_t)h.pos_begin;
#ifdef ROFL_EMUL
    // portable form (CUDA-on-CPU emulation): state in shared memory, one lane per Keccak word with barriers between the steps
    uint8_t *st8 = (uint8_t *)h.st;
    for (uint32_t e = 0; e <= nfull; e++) {
        const uint32_t base = STROBE_R * e;
        for (uint32_t p = lane; p < STROBE_R; p += TS_THREADS) {
            const uint32_t A = base + p;
            if (A >= A0 && A < A_end) st8[p] ^= absorb_stream_byte(A, base, A0, pb0, label, msgs);
        }
        TS_WSYNC();
        if (e < nfull) {                      // run_f closing block e: pos = 166
            if (lane == 0) { st8[STROBE_R] ^= absorb_close_pb(base, A0); st8[STROBE_R + 1] ^= 0x04 ^ 0x80; }
            TS_WSYNC();
            keccak_coop(h, lane);
        }
    }
#else
    // device form: lane t < 25 keeps word t of the state in a register and the permutation exchanges words with warp shuffles (keccak_shfl).
    // The absorbed byte stream is composed in shared memory 32 messages at a time -- one lane per message: its 9 framing bytes and 32 message
    // bytes (two 16-byte loads) go to a ring indexed by the absolute stream position -- so that a block costs every lane two aligned 8-byte reads
    // and a funnel shift.  (Composing each lane's 8 bytes of every block separately, 8 divisions and data-dependent branches per lane and block,
    // took as long as the permutation: 6 us per block in total.)
    uint64_t a = lane < 25 ? h.st[lane] : 0;
    uint8_t *ring8 = (uint8_t *)h.ring;
    const bool al16 = (((uintptr_t)msgs) & 15) == 0;
    for (uint32_t i = lane; i < TS_RING / 8; i += TS_THREADS) h.ring[i] = 0;           // positions before A0 (block 0) read as zero
    __syncwarp();
    uint32_t jnext = 0, filled = A0;                                                     // stream composed up to absolute position `filled`
    for (uint32_t e = 0; e <= nfull; e++) {
        const uint32_t base = STROBE_R * e, need = base + STROBE_R < A_end ? base + STROBE_R : A_end;
        while (filled < need) {                                                          // (at most 165 composed bytes are still unread here)
            const uint32_t j = jnext + lane;
            if (j < m) {
                const uint32_t a1 = A0 + 41u * j, b1 = a1 / STROBE_R * STROBE_R, b7 = (a1 + 7) / STROBE_R * STROBE_R, ap = a1 - 34;
                uint8_t hdr[9] = {(uint8_t)(j == 0 ? pb0 : (ap >= b1 ? ap - b1 + 1 : 0)), 0x12, label, 32, 0, 0, 0, (uint8_t)(a1 >= b7 ? a1 - b7 + 1 : 0), 0x02};
                for (int i = 0; i < 9; i++) ring8[(a1 + i) & (TS_RING - 1)] = hdr[i];
                const uint8_t *mp = msgs + 32 * (size_t)j;
                if (al16) {
                    const uint4 q0 = ((const uint4 *)mp)[0], q1 = ((const uint4 *)mp)[1];
                    const uint32_t wv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                    for (int i = 0; i < 32; i++) ring8[(a1 + 9 + i) & (TS_RING - 1)] = (uint8_t)(wv[i >> 2] >> (8 * (i & 3)));
                } else
                    for (int i = 0; i < 32; i++) ring8[(a1 + 9 + i) & (TS_RING - 1)] = mp[i];
            }
            jnext += TS_THREADS;
            filled = jnext < m ? A0 + 41u * jnext : A_end;
            if (jnext >= m) for (uint32_t i = lane; i < STROBE_R + 16; i += TS_THREADS) ring8[(A_end + i) & (TS_RING - 1)] = 0;      // the last block reads past the end
            __syncwarp();
        }
        if (lane < 21) {
            const uint32_t idx = (base + 8 * lane) & (TS_RING - 1), sh = (idx & 7) * 8;
            const uint64_t lo = h.ring[idx >> 3], hi = h.ring[((idx >> 3) + 1) & (TS_RING / 8 - 1)];
            uint64_t w = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
            if (lane == 20) {
                w &= 0x0000ffffffffffffULL;                                                // bytes 166, 167 of the block are not rate bytes
                if (e < nfull) w ^= ((uint64_t)absorb_close_pb(base, A0) << 48) | ((uint64_t)(0x04 ^ 0x80) << 56);
            }
            a ^= w;
        }
        if (e < nfull) a = keccak_shfl(a, lane);
        __syncwarp();                                                                    // every lane has read block e before the ring is written again
    }
    if (lane < 25) h.st[lane] = a;
#endif
    if (lane == 0) {
        const uint32_t a2l = A0 + 41u * (m - 1) + 7, cur = A_end / STROBE_R * STROBE_R;
        h.pos = A_end - cur;
        h.pos_begin = a2l >= cur ? a2l - cur + 1 : 0;
    }
    TS_WSYNC();
}
DEV void wt_absorb_many(strobe_sh &h, wts &w, int lane, uint8_t label, const uint8_t *msgs, uint32_t m) {
    if (m == 0) return;
    TS_WSYNC();
    if (lane == 0) { h.pos = w.pos; h.pos_begin = w.pos_begin; }
    TS_WSYNC();
    strobe_absorb_many_coop(h, lane, label, msgs, m);
    w.pos = h.pos; w.pos_begin = h.pos_begin; w.cur_flags = 2;
}
DEV bool is_zero32_(const uint8_t *b) { uint8_t z = 0; for (int i = 0; i < 32; i++) z |= b[i]; return z == 0; }
DEV void ts_challenge_sc(transcript &t, const char *label, sc &out) { uint8_t b[64]; transcript_challenge(t, label, b, 64); sc_from_bytes_wide(out, b); }
DEV void ts_append32(transcript &t, const char *label, const uint8_t *p) { transcript_append(t, label, p, 32); }
#endif

