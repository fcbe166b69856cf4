// Single-warp LATENCY of the serial sections of the Fiat-Shamir chain (cycles per operation, clock64): one active lane of one warp vs a
// full warp, for chains of fe_mul / fe_sq / ge_p3_dbl / ge_add / sc_mul and for sc_invert_vartime, fe_invsqrt-based ge_compress.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rofl-project-code_b200/csrc -I include tools/microbench_lat.cu -o tools/microbench_lat
#include <cstdio>
#include <cuda_runtime.h>
#define KG_TABLES 1
#include "kernels.cuh"
void rt_count_launch(const char *) {}
void *rt_prof_begin(int, cudaStream_t) { return nullptr; }
void rt_prof_end(int, void *, cudaStream_t) {}
void *rt_timeline_begin(const char *, cudaStream_t, unsigned) { return nullptr; }
void rt_timeline_end(void *, cudaStream_t) {}
#define N 256
__global__ void k_lat(long long *out, uint32_t *sink, int active) {
    if ((int)threadIdx.x >= active) return;
    fe x, y; for (int i = 0; i < 8; i++) { x.v[i] = threadIdx.x * 8 + i + 1; y.v[i] = 0x9e3779b9u * (i + 3); }
    long long t0 = clock64();
    for (int i = 0; i < N; i++) fe_mul(x, x, y);
    long long t1 = clock64();
    for (int i = 0; i < N; i++) fe_sq(x, x);
    long long t2 = clock64();
    ge_p3 P; ge_base(P); P.X.v[0] ^= x.v[0] & 1;
    for (int i = 0; i < N; i++) ge_p3_dbl(P, P);
    long long t3 = clock64();
    ge_p3 Q; ge_base(Q); 
    for (int i = 0; i < N; i++) ge_add(P, P, Q);
    long long t4 = clock64();
    sc s, t; for (int i = 0; i < 8; i++) { s.v[i] = x.v[i] >> 4; t.v[i] = y.v[i] >> 4; }
    for (int i = 0; i < N; i++) sc_mul(s, s, t);
    long long t5 = clock64();
    sc inv; for (int i = 0; i < 8; i++) sc_invert_vartime(inv, s), s.v[0] ^= inv.v[1] & 0xff;
    long long t6 = clock64();
    uint8_t enc[32]; for (int i = 0; i < 8; i++) { ge_compress(enc, P); P.X.v[0] ^= enc[3] & 1; }
    long long t7 = clock64();
    transcript tr; transcript_init(tr, "bench"); uint8_t b[64];
    for (int i = 0; i < 16; i++) { transcript_append(tr, "L", enc, 32); transcript_challenge(tr, "u", b, 64); enc[0] ^= b[0]; }
    long long t8 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / N; out[1] = (t2 - t1) / N; out[2] = (t3 - t2) / N; out[3] = (t4 - t3) / N; out[4] = (t5 - t4) / N; out[5] = (t6 - t5) / 8; out[6] = (t7 - t6) / 8; out[7] = (t8 - t7) / 16; }
    sink[threadIdx.x] = x.v[0] ^ P.T.v[3] ^ s.v[2] ^ inv.v[0] ^ enc[5] ^ b[1];
}
int main() {
    long long *d; uint32_t *sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 4096);
    const char *nm[8] = {"fe_mul", "fe_sq", "ge_p3_dbl", "ge_add(p3,p3)", "sc_mul", "sc_invert_vartime", "ge_compress", "append32+challenge64"};
    for (int active : {1, 32}) {
        k_lat<<<1, 32>>>(d, sink, active); cudaDeviceSynchronize();
        k_lat<<<1, 32>>>(d, sink, active); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("active lanes %d (cycles per operation, one warp alone on its SM):\n", active);
        for (int i = 0; i < 8; i++) printf("  %-22s %lld\n", nm[i], h[i]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
