// Device-side Fiat-Shamir: the Merlin / STROBE-128 transcripts of the Bulletproofs range proofs run on the GPU, so a whole
// RangeProof::prove_multiple / verify_multiple (bulletproofs 4.0.0; call sites rofl_crypto/src/range_proof_vec/mod.rs:124-135,
// 200-209) is ONE stream of kernel launches with no host round trip between its challenges.
//   k_ts_absorbV : Transcript::new(label); rangeproof_domain_sep(n, m); append_point("V", V_j) for the m commitments of a chunk.
//                  One warp per chunk: the 41*m stream bytes of a block of the sponge are generated in parallel (their STROBE
//                  framing bytes are a closed form of the stream position) and Keccak-f[1600] runs with one lane per 64-bit word.
//   k_ts_yz      : append A, S -> y, z (+ y^-1 and the 2^b-th power tables the polynomial kernels index)
//   k_ts_t12     : t1 = <l0+l1, r0+r1> - t0 - t2
//   k_ts_x       : append T_1, T_2 -> x;  t_x, t_x_blinding, e_blinding -> w;  innerproduct_domain_sep(N)
//   k_ts_round   : append L, R -> u;  u^2, u^-2, u^-2 y^-n', running products, the coefficient tables of the unfolded / frozen
//                  rounds and the digit strings the next generator kernel consumes
//   k_ts_final   : a = a^ prod u_k, b = b^ prod u_k^-1 (proofs without the fused tail kernel)
//   k_ts_verify / k_verify_keys / k_verify_prep : the verifier's replay, the Fiat-Shamir derivation of its batching scalars over ALL
//                  chunks of the call, and the per-chunk scalar block of the batched check (kernels.cuh K7)
// Included by kernels.cuh (kernel group KG_TS).
#pragma once

// ---- prover ----------------------------------------------------------------------------------------------------------------
struct ts_absorb_args { transcript *ts; const uint8_t *V32; uint32_t m, n; int label_id; };       // label_id 0 "RangeProof", 1 "L2RangeProof"
struct ts_yz_args { transcript *ts; const uint8_t *AS; uint8_t *proofs; uint32_t plen, C; sc_st *ypow2, *zpow2, *yinvpow2, *z; };
struct ts_x_args { transcript *ts; const uint8_t *T12; uint8_t *proofs; uint32_t plen, C; const sc_st *tsum, *sums; sc_st *x, *w2; uint64_t N; };
struct ts_round_args {
    transcript *ts; const uint8_t *LR; uint8_t *proofs; uint32_t plen, off, C;
    const sc_st *yinvpow2; int lgnp;
    sc_st *u2, *uinv2, *up;                        // [C], [C], [2C] (prod u | prod u^-1)
    sc_st *coefG, *coefH; uint32_t cstride, nblk;  // coefficient tables (nullable): c'[2t] = c[t], c'[2t+1] = c[t] s when 2 nblk <= cstride
    int emit;                                      // 0 nothing, 1 width-FOLD_W NAFs of (u^2, u^-2 y^-n'), 2 table digits of the coefficients, 3 radix-16 digits of them, 4 width-3 NAFs of them
    uint32_t rtK[9]; int rtc, rtnw;
    int8_t *nafs; int16_t *digs16; int8_t *digs8;
};
struct ts_final_args { const sc_st *a, *b, *up; uint8_t *proofs; uint32_t plen, off, C; size_t N; };
#ifdef KG_TS
KERNEL void LB(TS_THREADS, 1) k_ts_absorbV(ts_absorb_args a) {
    __shared__ strobe_sh h;
    const int c = blockIdx.x, lane = threadIdx.x;
    wts w;
    wt_init(h, w, lane, a.label_id ? "L2RangeProof" : "RangeProof");
    wt_append(h, w, lane, "dom-sep", (const uint8_t *)"rangeproof v1", 13);
    wt_append_u64(h, w, lane, "n", (uint64_t)a.n); wt_append_u64(h, w, lane, "m", (uint64_t)a.m);
    wt_absorb_many(h, w, lane, (uint8_t)'V', a.V32 + 32 * (size_t)c * a.m, a.m);
    wt_store(h, w, lane, a.ts[c]);
}
KLAUNCH(k_ts_absorbV, true, (ts_absorb_args a), (a))
KERNEL void LB(TS_THREADS, 1) k_ts_yz(ts_yz_args a) {
    __shared__ strobe_sh h;
    const uint32_t c = blockIdx.x; const int lane = threadIdx.x;
    wts w; wt_load(h, w, lane, a.ts[c]);
    uint8_t *o = a.proofs + (size_t)a.plen * c;
    wt_load32(h, lane, a.AS + 32 * (size_t)c); o[lane] = h.io[lane]; wt_append(h, w, lane, "A", h.io, 32);
    wt_load32(h, lane, a.AS + 32 * (size_t)(a.C + c)); o[32 + lane] = h.io[lane]; wt_append(h, w, lane, "S", h.io, 32);
    sc y, z; wt_challenge_sc(h, w, lane, "y", y); wt_challenge_sc(h, w, lane, "z", z);
    wt_store(h, w, lane, a.ts[c]);
    if (lane == 1) st_sc(a.z + c, z);
    if (lane < 3) {                                                // y^(2^b), z^(2^b), y^-(2^b)
        sc cur = lane == 1 ? z : y;
        if (lane == 2) sc_invert_vartime(cur, y);
        sc_st *tab = (lane == 0 ? a.ypow2 : lane == 1 ? a.zpow2 : a.yinvpow2) + 32 * (size_t)c;
        for (int b = 0; b < 32; b++) { st_sc(tab + b, cur); sc_mul(cur, cur, cur); }
    }
}
KLAUNCH(k_ts_yz, true, (ts_yz_args a), (a))
// t12[c] = t1 = tsum[C + c] - t0 - t2, t12[C + c] = t2     (tsum = t0 | <l0+l1, r0+r1> | t2, k_poly / k_sc_sum)
KERNEL void k_ts_t12(sc_st *t12, const sc_st *tsum, uint32_t C) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    sc t0, tt, t2; ld_sc(t0, tsum + c); ld_sc(tt, tsum + C + c); ld_sc(t2, tsum + 2 * C + c);
    sc_sub(tt, tt, t0); sc_sub(tt, tt, t2);
    st_sc(t12 + c, tt); st_sc(t12 + C + c, t2);
}
KLAUNCH(k_ts_t12, false, (sc_st *t12, const sc_st *tsum, uint32_t C), (t12, tsum, C))
KERNEL void LB(TS_THREADS, 1) k_ts_x(ts_x_args a) {
    __shared__ strobe_sh h;
    __shared__ uint8_t sb[96];                                     // t_x | t_x_blinding | e_blinding
    const uint32_t c = blockIdx.x, C = a.C; const int lane = threadIdx.x;
    wts w; wt_load(h, w, lane, a.ts[c]);
    uint8_t *o = a.proofs + (size_t)a.plen * c;
    wt_load32(h, lane, a.T12 + 32 * (size_t)c); o[64 + lane] = h.io[lane]; wt_append(h, w, lane, "T_1", h.io, 32);
    wt_load32(h, lane, a.T12 + 32 * (size_t)(C + c)); o[96 + lane] = h.io[lane]; wt_append(h, w, lane, "T_2", h.io, 32);
    sc x, wch; wt_challenge_sc(h, w, lane, "x", x);
    if (lane == 0) {
        sc t0, t1, t2, sa, ss, st1, st2, szg, xx, tx, txb, eb, tmp;
        ld_sc(t0, a.tsum + c); ld_sc(t1, a.tsum + C + c); ld_sc(t2, a.tsum + 2 * C + c); sc_sub(t1, t1, t0); sc_sub(t1, t1, t2);
        ld_sc(sa, a.sums + c); ld_sc(ss, a.sums + C + c); ld_sc(st1, a.sums + 2 * C + c); ld_sc(st2, a.sums + 3 * C + c); ld_sc(szg, a.sums + 4 * C + c);
        sc_mul(xx, x, x);
        sc_mul(tmp, t1, x); sc_add(tx, t0, tmp); sc_mul(tmp, t2, xx); sc_add(tx, tx, tmp);
        sc_mul(tmp, st1, x); sc_add(txb, szg, tmp); sc_mul(tmp, st2, xx); sc_add(txb, txb, tmp);
        sc_mul(tmp, ss, x); sc_add(eb, sa, tmp);
        uint8_t b[32];
        sc_tobytes(b, tx); for (int i = 0; i < 32; i++) sb[i] = b[i];
        sc_tobytes(b, txb); for (int i = 0; i < 32; i++) sb[32 + i] = b[i];
        sc_tobytes(b, eb); for (int i = 0; i < 32; i++) sb[64 + i] = b[i];
    }
    TS_WSYNC();
    for (int i = lane; i < 96; i += TS_THREADS) o[128 + i] = sb[i];
    wt_append(h, w, lane, "t_x", sb, 32); wt_append(h, w, lane, "t_x_blinding", sb + 32, 32); wt_append(h, w, lane, "e_blinding", sb + 64, 32);
    wt_challenge_sc(h, w, lane, "w", wch);
    wt_append(h, w, lane, "dom-sep", (const uint8_t *)"ipp v1", 6); wt_append_u64(h, w, lane, "n", a.N);
    wt_store(h, w, lane, a.ts[c]);
    if (lane == 0) { st_sc(a.x + c, x); st_sc(a.w2 + c, wch); st_sc(a.w2 + C + c, wch); }
}
KLAUNCH(k_ts_x, true, (ts_x_args a), (a))
KERNEL void LB(TS_THREADS, 1) k_ts_round(ts_round_args a) {
    __shared__ sc_st sf[2];                                       // u^2 | u^-2 y^-n'
    __shared__ strobe_sh h;
    const uint32_t c = blockIdx.x, C = a.C; const int lane = threadIdx.x;
    wts w; wt_load(h, w, lane, a.ts[c]);
    uint8_t *o = a.proofs + (size_t)a.plen * c + a.off;
    wt_load32(h, lane, a.LR + 32 * (size_t)c); o[lane] = h.io[lane]; wt_append(h, w, lane, "L", h.io, 32);
    wt_load32(h, lane, a.LR + 32 * (size_t)(C + c)); o[32 + lane] = h.io[lane]; wt_append(h, w, lane, "R", h.io, 32);
    sc u; wt_challenge_sc(h, w, lane, "u", u);
    wt_store(h, w, lane, a.ts[c]);
    if (lane == 0) {
        sc ui, u2, ui2, sH, yp, p; sc_invert_vartime(ui, u);
        sc_mul(u2, u, u); sc_mul(ui2, ui, ui); ld_sc(yp, a.yinvpow2 + 32 * (size_t)c + a.lgnp); sc_mul(sH, ui2, yp);
        st_sc(a.u2 + c, u2); st_sc(a.uinv2 + c, ui2); st_sc(sf, u2); st_sc(sf + 1, sH);
        ld_sc(p, a.up + c); sc_mul(p, p, u); st_sc(a.up + c, p);
        ld_sc(p, a.up + C + c); sc_mul(p, p, ui); st_sc(a.up + C + c, p);
        if (a.emit == 1) { sc_naf(a.nafs + ((size_t)c * 2) * 256, u2, FOLD_W); sc_naf(a.nafs + ((size_t)c * 2 + 1) * 256, sH, FOLD_W); }
    }
    TS_WSYNC();
    uint32_t nout = a.nblk;
    if (a.coefG && 2 * a.nblk <= a.cstride) {
        sc_st *g0 = a.coefG + (size_t)c * a.cstride, *h0 = a.coefH + (size_t)c * a.cstride;
        // in place, highest block first inside a pass of 32: every lane reads its entries before any lane writes
        for (uint32_t base = (a.nblk - 1) / TS_THREADS * TS_THREADS;; base -= TS_THREADS) {
            const uint32_t t = base + lane; sc gt, ht;
            if (t < a.nblk) { ld_sc(gt, g0 + t); ld_sc(ht, h0 + t); }
            TS_WSYNC();
            if (t < a.nblk) { sc f, x; st_sc(g0 + 2 * t, gt); ld_sc(f, sf); sc_mul(x, gt, f); st_sc(g0 + 2 * t + 1, x); st_sc(h0 + 2 * t, ht); ld_sc(f, sf + 1); sc_mul(x, ht, f); st_sc(h0 + 2 * t + 1, x); }
            TS_WSYNC();
            if (base == 0) break;
        }
        nout = 2 * a.nblk;
    }
    if (a.coefG && a.emit >= 2) {
        for (uint32_t t = lane; t < 2 * nout; t += TS_THREADS) {
            const uint32_t which = t / nout, i = t % nout;
            sc v; ld_sc(v, (which ? a.coefH : a.coefG) + (size_t)c * a.cstride + i);
            if (a.emit == 2) {
                uint32_t xk[9]; msm_recode(xk, v, a.rtK); int16_t *d = a.digs16 + (((size_t)c * 2 + which) * nout + i) * RT_MAXW;
                for (int w = 0; w < a.rtnw; w++) d[w] = (int16_t)msm_digit(xk, w, a.rtc);
            } else if (a.emit == 3) sc_radix16(a.digs8 + (((size_t)c * 2 + which) * nout + i) * 64, v);
            else sc_naf(a.nafs + (((size_t)c * 2 + which) * nout + i) * 256, v, 3);
        }
    }
}
KLAUNCH(k_ts_round, true, (ts_round_args a), (a))
KERNEL void k_ts_final(ts_final_args a) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    sc x, y, p; ld_sc(x, a.a + (size_t)c * a.N); ld_sc(y, a.b + (size_t)c * a.N);
    ld_sc(p, a.up + c); sc_mul(x, x, p); ld_sc(p, a.up + a.C + c); sc_mul(y, y, p);
    uint8_t *o = a.proofs + (size_t)a.plen * c + a.off; uint8_t e[32];
    sc_tobytes(e, x); st_bytes32(o, e); sc_tobytes(e, y); st_bytes32(o + 32, e);
}
KLAUNCH(k_ts_final, false, (ts_final_args a), (a))
#endif
void launch_k_ts_absorbV(dim3 g_, dim3 b_, cudaStream_t s_, ts_absorb_args a);
void launch_k_ts_yz(dim3 g_, dim3 b_, cudaStream_t s_, ts_yz_args a);
void launch_k_ts_t12(dim3 g_, dim3 b_, cudaStream_t s_, sc_st *t12, const sc_st *tsum, uint32_t C);
void launch_k_ts_x(dim3 g_, dim3 b_, cudaStream_t s_, ts_x_args a);
void launch_k_ts_round(dim3 g_, dim3 b_, cudaStream_t s_, ts_round_args a);
void launch_k_ts_final(dim3 g_, dim3 b_, cudaStream_t s_, ts_final_args a);

// ---- verifier --------------------------------------------------------------------------------------------------------------
// k_ts_verify: replay of RangeProof::verify_multiple's transcript for every chunk: challenges chal[c] = y z x w u_0..u_(lgN-1),
//   bad[c] |= 1 when A, S, T_1, T_2, L_k or R_k is the identity encoding (validate_and_append_point), the chunk's small points
//   sp32[c] = A S T1 T2 L.. R.. H B, and digest[c] = 32 challenge bytes drawn after a and b have been absorbed as well: the digest binds
//   EVERY byte of the chunk's proof and commitments.
// k_verify_keys: the batching scalars.  bulletproofs draws its per-proof scalar c from transcript.build_rng().finalize(thread_rng); here
//   (c_i, rho_i) = ChaCha20(SHA3-256(SHA3-256(seed | C | digest_0 .. digest_(C-1)) | domain | chunk index)): a function of the verifier's
//   seed AND of all proofs of the call, so that the random linear combination over chunks stays sound even for a known seed.
// k_verify_prep: per-chunk scalar block of the batched check (layout: kernels.cuh K7)
struct ts_verify_args { const transcript *ts; const uint8_t *proofs; uint32_t plen, C; int lgN; uint64_t N; sc_st *chal; uint8_t *digest; int *bad; uint8_t *sp32; uint8_t B32[32], H32[32]; };
struct verify_keys_args { const uint8_t *digest; uint32_t C, dom; uint64_t c_off; uint8_t seed[32]; sc_st *ccrho; };       // ccrho[c] = c_c, ccrho[C + c] = rho_c
struct verify_prep_args {
    const uint8_t *proofs; uint32_t plen, C; int lgN, lgm, n; const sc_st *vch, *ccrho;
    sc_st *chal; int chs; sc_st *small; int nsmall; sc_st *yinvpow2, *zpow2;
};
#ifdef KG_TS
KERNEL void LB(TS_THREADS, 1) k_ts_verify(ts_verify_args a) {
    __shared__ strobe_sh h;
    const uint32_t c = blockIdx.x; const int lane = threadIdx.x;          // one warp per chunk
    const int lgN = a.lgN, nsmall = 6 + 2 * lgN;
    const uint8_t *p = a.proofs + (size_t)a.plen * c, *ipp = p + 224;
    uint8_t *sp = a.sp32 + 32 * (size_t)c * nsmall;
    sc_st *ch = a.chal + (size_t)c * (4 + lgN);
    wts w; wt_load(h, w, lane, a.ts[c]);
    int ok = 1; sc s;
    // validate_and_append_point: the identity encoding is refused
    ok &= !wt_load32(h, lane, p); sp[lane] = h.io[lane]; wt_append(h, w, lane, "A", h.io, 32);
    ok &= !wt_load32(h, lane, p + 32); sp[32 + lane] = h.io[lane]; wt_append(h, w, lane, "S", h.io, 32);
    wt_challenge_sc(h, w, lane, "y", s); if (lane == 0) st_sc(ch, s);
    wt_challenge_sc(h, w, lane, "z", s); if (lane == 0) st_sc(ch + 1, s);
    ok &= !wt_load32(h, lane, p + 64); sp[64 + lane] = h.io[lane]; wt_append(h, w, lane, "T_1", h.io, 32);
    ok &= !wt_load32(h, lane, p + 96); sp[96 + lane] = h.io[lane]; wt_append(h, w, lane, "T_2", h.io, 32);
    wt_challenge_sc(h, w, lane, "x", s); if (lane == 0) st_sc(ch + 2, s);
    wt_load32(h, lane, p + 128); wt_append(h, w, lane, "t_x", h.io, 32);
    wt_load32(h, lane, p + 160); wt_append(h, w, lane, "t_x_blinding", h.io, 32);
    wt_load32(h, lane, p + 192); wt_append(h, w, lane, "e_blinding", h.io, 32);
    wt_challenge_sc(h, w, lane, "w", s); if (lane == 0) st_sc(ch + 3, s);
    wt_append(h, w, lane, "dom-sep", (const uint8_t *)"ipp v1", 6); wt_append_u64(h, w, lane, "n", a.N);
    for (int k = 0; k < lgN; k++) {
        ok &= !wt_load32(h, lane, ipp + 64 * k); sp[32 * (4 + k) + lane] = h.io[lane]; wt_append(h, w, lane, "L", h.io, 32);
        ok &= !wt_load32(h, lane, ipp + 64 * k + 32); sp[32 * (4 + lgN + k) + lane] = h.io[lane]; wt_append(h, w, lane, "R", h.io, 32);
        wt_challenge_sc(h, w, lane, "u", s); if (lane == 0) st_sc(ch + 4 + k, s);
    }
    sp[32 * (4 + 2 * lgN) + lane] = a.H32[lane]; sp[32 * (5 + 2 * lgN) + lane] = a.B32[lane];
    wt_load32(h, lane, ipp + 64 * lgN); wt_append(h, w, lane, "a", h.io, 32);
    wt_load32(h, lane, ipp + 64 * lgN + 32); wt_append(h, w, lane, "b", h.io, 32);
    wt_challenge(h, w, lane, "rofl-batch", h.io, 32);
    a.digest[32 * (size_t)c + lane] = h.io[lane];
    if (!ok && lane == 0) a.bad[c] = 1;
}
KLAUNCH(k_ts_verify, true, (ts_verify_args a), (a))
KERNEL void LB(256, 1) k_verify_keys(verify_keys_args a) {
    __shared__ strobe_sh h;                                       // (used as a plain SHA3-256 sponge here: rate 136)
    __shared__ uint8_t glob[32];
    // glob = SHA3-256(seed | C | digest_0 .. digest_(C-1)): warp 0 absorbs the bytes side by side and permutes with shuffles.  (Control flow is
    // uniform over the block -- the other warps walk through the same barriers and permute nothing -- so that the CPU emulation, whose warp
    // barrier is the block barrier, runs the same code.)
    const int lane = threadIdx.x; const bool w0 = threadIdx.x < TS_THREADS;
    uint8_t *st8 = (uint8_t *)h.st;
    if (lane < 25) h.st[lane] = 0;
    TS_WSYNC();
    const uint64_t total = 40 + 32 * (uint64_t)a.C;
    for (uint64_t base = 0;; base += 136) {
        const uint64_t left = total - base; const uint32_t n = left < 136 ? (uint32_t)left : 136;
        if (w0) for (uint32_t i = lane; i < n; i += TS_THREADS) {
            const uint64_t k = base + i;
            st8[i] ^= k < 32 ? a.seed[k] : k < 40 ? (uint8_t)((uint64_t)a.C >> (8 * (k - 32))) : a.digest[k - 40];
        }
        TS_WSYNC();
        if (n < 136) { if (lane == 0) { st8[n] ^= 0x06; st8[135] ^= 0x80; } TS_WSYNC(); wt_permute(h, lane); break; }
        wt_permute(h, lane);
    }
    if (w0) glob[lane] = st8[lane];
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < a.C; c += blockDim.x) {
        uint8_t buf[44], key[32]; for (int i = 0; i < 32; i++) buf[i] = glob[i];
        const uint64_t idx = a.c_off + c;
        for (int i = 0; i < 4; i++) buf[32 + i] = (uint8_t)(a.dom >> (8 * i));
        for (int i = 0; i < 8; i++) buf[36 + i] = (uint8_t)(idx >> (8 * i));
        sha3_256(key, buf, 44);
        uint32_t kw[8]; for (int i = 0; i < 8; i++) kw[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
        sc v; nonce_scalar(v, kw, 0); st_sc(a.ccrho + c, v); nonce_scalar(v, kw, 1); st_sc(a.ccrho + a.C + c, v);
    }
}
KLAUNCH(k_verify_keys, true, (verify_keys_args a), (a))
KERNEL void LB(2 * TS_THREADS, 1) k_verify_prep(verify_prep_args a) {
    // two warps per chunk: thread 0 inverts y and the u_k (Montgomery's trick, one inversion) while warp 1 computes what needs no inverse
    // (thread 32: the constant block and delta, thread 33: the powers z^(2^b)); different paths of ONE warp would run one after the other
    __shared__ sc_st inv[34];                                     // y^-1, u_k^-1
    const uint32_t c = blockIdx.x, C = a.C; const int tid = threadIdx.x, lgN = a.lgN;
    const sc_st *vch = a.vch + (size_t)c * (4 + lgN);
    sc rho, cc; ld_sc(cc, a.ccrho + c); ld_sc(rho, a.ccrho + C + c);
    sc_st *ch = a.chal + (size_t)c * a.chs, *sm = a.small + (size_t)c * a.nsmall;
    if (tid == 0) {                                              // (zero, probability 2^-252, is replaced by 1)
        sc pre[34], v[34], acc, one; sc_from_u64(one, 1); acc = one;
        for (int i = 0; i <= lgN; i++) { ld_sc(v[i], i == 0 ? vch : vch + 3 + i); if (sc_iszero(v[i])) v[i] = one; pre[i] = acc; sc_mul(acc, acc, v[i]); }
        sc iv; sc_invert_vartime(iv, acc);
        for (int i = lgN; i >= 0; i--) { sc t; sc_mul(t, iv, pre[i]); sc_mul(iv, iv, v[i]); st_sc(inv + i, t); }
    }
    if (tid == TS_THREADS + 1) {
        sc cur; ld_sc(cur, vch + 1);
        sc_st *tab = a.zpow2 + 32 * (size_t)c;
        for (int b = 0; b < 32; b++) { st_sc(tab + b, cur); sc_mul(cur, cur, cur); }
    }
    if (tid == TS_THREADS) {
        const uint8_t *p = a.proofs + (size_t)a.plen * c, *ipp = p + 224; uint8_t b32[32];
        sc y, z, x, w, t_x, t_xb, e_bl, pa, pb, zz, tmp, tmp2, cx;
        ld_sc(y, vch); ld_sc(z, vch + 1); ld_sc(x, vch + 2); ld_sc(w, vch + 3);
        ld_bytes32(b32, p + 128); sc_frombytes(t_x, b32); ld_bytes32(b32, p + 160); sc_frombytes(t_xb, b32); ld_bytes32(b32, p + 192); sc_frombytes(e_bl, b32);
        ld_bytes32(b32, ipp + 64 * lgN); sc_frombytes(pa, b32); ld_bytes32(b32, ipp + 64 * lgN + 32); sc_frombytes(pb, b32);
        sc_mul(zz, z, z);
        sc_mul(tmp, rho, z); st_sc(ch, tmp); sc_mul(tmp, rho, zz); st_sc(ch + 1, tmp);
        sc_mul(tmp, rho, pa); st_sc(ch + 2, tmp); sc_mul(tmp, rho, pb); st_sc(ch + 3, tmp); st_sc(ch + 4, cc);
        sc_mul(cx, cc, x);
        st_sc(sm, rho); sc_mul(tmp, x, rho); st_sc(sm + 1, tmp); sc_mul(tmp, cx, rho); st_sc(sm + 2, tmp); sc_mul(tmp, cx, x); sc_mul(tmp, tmp, rho); st_sc(sm + 3, tmp);
        sc_mul(tmp, cc, t_xb); sc_add(tmp, tmp, e_bl); sc_neg(tmp, tmp); sc_mul(tmp, tmp, rho); st_sc(sm + 4 + 2 * lgN, tmp);       // H: -e_bl - c t_x_bl
        // delta = (z - zz) sum_{i<N} y^i - z^3 (2^n - 1) sum_{j<m} z^j ; sums of powers via prod (1 + s^(2^b))
        sc one, sum_y, sum_z, pw, delta, s2; sc_from_u64(one, 1); sum_y = one; sum_z = one;
        pw = y; for (int bb = 0; bb < lgN; bb++) { sc_add(tmp, one, pw); sc_mul(sum_y, sum_y, tmp); sc_mul(pw, pw, pw); }
        pw = z; for (int bb = 0; bb < a.lgm; bb++) { sc_add(tmp, one, pw); sc_mul(sum_z, sum_z, tmp); sc_mul(pw, pw, pw); }
        sc_from_u64(s2, a.n == 64 ? ~0ULL : ((1ULL << a.n) - 1));
        sc_sub(delta, z, zz); sc_mul(delta, delta, sum_y);
        sc_mul(tmp, zz, z); sc_mul(tmp, tmp, s2); sc_mul(tmp, tmp, sum_z); sc_sub(delta, delta, tmp);
        sc_mul(tmp, pa, pb); sc_sub(tmp, t_x, tmp); sc_mul(tmp, w, tmp);                                              // w (t_x - a b)
        sc_sub(tmp2, delta, t_x); sc_mul(tmp2, cc, tmp2); sc_add(tmp, tmp, tmp2); sc_mul(tmp, tmp, rho); st_sc(sm + 5 + 2 * lgN, tmp);      // B
    }
    __syncthreads();
    if (tid == 1) {
        sc cur; ld_sc(cur, inv);
        sc_st *tab = a.yinvpow2 + 32 * (size_t)c;
        for (int b = 0; b < 32; b++) { st_sc(tab + b, cur); sc_mul(cur, cur, cur); }
    }
    for (int k = tid - TS_THREADS; k >= 0 && k < lgN; k += TS_THREADS) {        // warp 1, beside the power table of warp 0
        sc u, ui, t; ld_sc(u, vch + 4 + k); ld_sc(ui, inv + 1 + k);
        st_sc(ch + 5 + k, u); st_sc(ch + 5 + lgN + k, ui);
        sc_mul(t, u, u); sc_mul(t, t, rho); st_sc(sm + 4 + k, t);
        sc_mul(t, ui, ui); sc_mul(t, t, rho); st_sc(sm + 4 + lgN + k, t);
    }
}
KLAUNCH(k_verify_prep, true, (verify_prep_args a), (a))
#endif
void launch_k_ts_verify(dim3 g_, dim3 b_, cudaStream_t s_, ts_verify_args a);
void launch_k_verify_keys(dim3 g_, dim3 b_, cudaStream_t s_, verify_keys_args a);
void launch_k_verify_prep(dim3 g_, dim3 b_, cudaStream_t s_, verify_prep_args a);
