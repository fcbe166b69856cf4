// Debug tool: runs a few point-arithmetic patterns on the GPU and prints their compressed results, so that builds with
// different compile options (e.g. -DFE_INLINE_MUL=1 vs 0) can be diffed.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../rofl-project-code_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#define KG_ALL 1
#include "kernels.cuh"
void rt_count_launch(const char *) {}
void *rt_prof_begin(int, cudaStream_t) { return nullptr; }
void rt_prof_end(int, void *, cudaStream_t) {}
__global__ void k_test(uint8_t *out, niels_st *scratch, p3_st *gB) {
    __shared__ p3_st buf[128];
    int tid = threadIdx.x;
    ge_p3 B; ge_base(B);
    ge_p3 acc; ge_p3_0(acc);
    // build niels of (tid+1)*B by repeated madd of B's niels
    fe zi; fe_invert(zi, B.Z); ge_niels nb; ge_p3_to_niels(nb, B, zi);
    st_niels(scratch + tid, nb);
    for (int i = 0; i <= (tid & 7); i++) acc_add_niels(acc, scratch + tid, (i & 1) && (tid & 8));
    uint8_t e[32];
    if (tid < 16) { ge_compress(e, acc); for (int k = 0; k < 32; k++) out[32 * tid + k] = e[k]; }
    block_sum_p3(acc, buf, tid, blockDim.x);
    if (tid == 0) { ge_compress(e, acc); for (int k = 0; k < 32; k++) out[32 * 16 + k] = e[k]; }
    // doubling chain + add
    ge_p3 d = B; for (int i = 0; i < 5; i++) ge_p3_dbl(d, d); ge_add(d, d, B);
    if (tid == 1) { ge_compress(e, d); for (int k = 0; k < 32; k++) out[32 * 17 + k] = e[k]; }
    if (tid == 2) {
        ge_p3 q = B;
        for (int i = 0; i < 5; i++) { ge_p3_dbl(q, q); ge_compress(e, q); for (int k = 0; k < 32; k++) out[32 * (18 + i) + k] = e[k]; }
        fe x, y, z; fe_add(x, B.X, B.Y); fe_sq(y, x); z = x; fe_sq(z, z);
        fe_tobytes(e, y); for (int k = 0; k < 32; k++) out[32 * 23 + k] = e[k];
        fe_tobytes(e, z); for (int k = 0; k < 32; k++) out[32 * 24 + k] = e[k];
        ge_p1p1 t; ge_dbl_p1p1(t, B.X, B.Y, B.Z);
        fe_tobytes(e, t.X); for (int k = 0; k < 32; k++) out[32 * 25 + k] = e[k];
        fe_tobytes(e, t.Y); for (int k = 0; k < 32; k++) out[32 * 26 + k] = e[k];
        fe_tobytes(e, t.Z); for (int k = 0; k < 32; k++) out[32 * 27 + k] = e[k];
        fe_tobytes(e, t.T); for (int k = 0; k < 32; k++) out[32 * 28 + k] = e[k];
        {   // hypotheses for the failing add
            ge_p3 Bg; ld_p3(Bg, gB);                    // base point through global memory (not a compile-time constant)
            ge_p3 r1, r2; ge_add(r1, q, B); ge_add(r2, q, Bg);
            ge_compress(e, r1); for (int k = 0; k < 32; k++) out[32 * 33 + k] = e[k];
            ge_compress(e, r2); for (int k = 0; k < 32; k++) out[32 * 34 + k] = e[k];
            ge_cached c1, c2; ge_p3_to_cached(c1, B); ge_p3_to_cached(c2, Bg);
            fe_tobytes(e, c1.T2d); for (int k = 0; k < 32; k++) out[32 * 35 + k] = e[k];
            fe_tobytes(e, c2.T2d); for (int k = 0; k < 32; k++) out[32 * 36 + k] = e[k];
            fe_tobytes(e, c1.YplusX); for (int k = 0; k < 32; k++) out[32 * 37 + k] = e[k];
            fe_tobytes(e, c2.YplusX); for (int k = 0; k < 32; k++) out[32 * 38 + k] = e[k];
            fe m1, m2; fe_mul(m1, GE_BASE_T, GE_D2); fe tt = Bg.T, dd; for (int k = 0; k < 8; k++) dd.v[k] = GE_D2.v[k] ^ (gB->w[0] & 0); fe_mul(m2, tt, dd);
            fe_tobytes(e, m1); for (int k = 0; k < 32; k++) out[32 * 39 + k] = e[k];
            fe_tobytes(e, m2); for (int k = 0; k < 32; k++) out[32 * 40 + k] = e[k];
        }
        ge_p3 r; ge_p1p1_to_p3(r, t);
        fe_tobytes(e, r.X); for (int k = 0; k < 32; k++) out[32 * 29 + k] = e[k];
        fe_tobytes(e, r.Y); for (int k = 0; k < 32; k++) out[32 * 30 + k] = e[k];
        fe_tobytes(e, r.Z); for (int k = 0; k < 32; k++) out[32 * 31 + k] = e[k];
        fe_tobytes(e, r.T); for (int k = 0; k < 32; k++) out[32 * 32 + k] = e[k];
    }
}
int main() {
    uint8_t *d_out; niels_st *d_s; cudaMalloc(&d_out, 32 * 41); cudaMalloc(&d_s, sizeof(niels_st) * 128);
    p3_st *d_B; cudaMalloc(&d_B, sizeof(p3_st)); { ge_p3 B; ge_base(B); p3_st hb; st_p3(&hb, B); cudaMemcpy(d_B, &hb, sizeof(hb), cudaMemcpyHostToDevice); }
    k_test<<<1, 128>>>(d_out, d_s, d_B);
    uint8_t h[32 * 41]; cudaError_t err = cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("err=%d\n", (int)err);
    for (int i = 0; i < 41; i++) { for (int k = 0; k < 32; k++) printf("%02x", h[32 * i + k]); printf("\n"); }
    return 0;
}
