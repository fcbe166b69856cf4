// Integer-pipe micro-benchmarks for the roofline denominator (SURVEY.md 8d asks for a measured IMAD.WIDE peak):
// independent mad.wide.u32 / mad.lo.u32 / DFMA chains in registers on all SMs, and dependent fe_mul / fe_sq chains.
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../rofl-project-code_b200/csrc/ge25519.cuh"

#define ITERS 4096
__global__ void k_imad_wide(uint64_t *out, uint32_t a, uint32_t b) {
    uint64_t c0 = threadIdx.x, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    uint32_t x = a + threadIdx.x, y = b;
    for (int i = 0; i < ITERS; i++) {
        asm volatile("mad.wide.u32 %0, %8, %9, %0;\n\tmad.wide.u32 %1, %8, %9, %1;\n\tmad.wide.u32 %2, %8, %9, %2;\n\tmad.wide.u32 %3, %8, %9, %3;\n\t"
                     "mad.wide.u32 %4, %8, %9, %4;\n\tmad.wide.u32 %5, %8, %9, %5;\n\tmad.wide.u32 %6, %8, %9, %6;\n\tmad.wide.u32 %7, %8, %9, %7;"
                     : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3), "+l"(c4), "+l"(c5), "+l"(c6), "+l"(c7) : "r"(x), "r"(y));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
__global__ void k_imad_lo(uint32_t *out, uint32_t a, uint32_t b) {
    uint32_t c0 = threadIdx.x, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    uint32_t x = a + threadIdx.x, y = b;
    for (int i = 0; i < ITERS; i++) {
        asm volatile("mad.lo.u32 %0, %8, %9, %0;\n\tmad.lo.u32 %1, %8, %9, %1;\n\tmad.lo.u32 %2, %8, %9, %2;\n\tmad.lo.u32 %3, %8, %9, %3;\n\t"
                     "mad.lo.u32 %4, %8, %9, %4;\n\tmad.lo.u32 %5, %8, %9, %5;\n\tmad.lo.u32 %6, %8, %9, %6;\n\tmad.lo.u32 %7, %8, %9, %7;"
                     : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7) : "r"(x), "r"(y));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
__global__ void k_dfma(double *out, double a, double b) {
    double c0 = threadIdx.x, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    for (int i = 0; i < ITERS; i++) {
        c0 = fma(a, b, c0); c1 = fma(a, b, c1); c2 = fma(a, b, c2); c3 = fma(a, b, c3);
        c4 = fma(a, b, c4); c5 = fma(a, b, c5); c6 = fma(a, b, c6); c7 = fma(a, b, c7);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
#define FE_ITERS 2048
__global__ void k_fe_mul(uint32_t *out, uint32_t seed) {
    fe a, b; for (int i = 0; i < 10; i++) { a.v[i] = (seed * (i + 1) + threadIdx.x) & 0x1ffffff; b.v[i] = (seed * (i + 7) + blockIdx.x) & 0x1ffffff; }
    for (int i = 0; i < FE_ITERS; i++) { fe_mul(a, a, b); fe_mul(b, b, a); }
    uint32_t r = 0; for (int i = 0; i < 10; i++) r ^= a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_fe_sq(uint32_t *out, uint32_t seed) {
    fe a, b; for (int i = 0; i < 10; i++) { a.v[i] = (seed * (i + 1) + threadIdx.x) & 0x1ffffff; b.v[i] = (seed * (i + 7) + blockIdx.x) & 0x1ffffff; }
    for (int i = 0; i < FE_ITERS; i++) { fe_sq(a, a); fe_sq(b, b); }
    uint32_t r = 0; for (int i = 0; i < 10; i++) r ^= a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// one point doubling chain (the fold kernel's inner loop): 4S + 3M + carry
__global__ void k_dbl(uint32_t *out, uint32_t seed) {
    ge_p3 p; ge_base(p); p.X.v[0] ^= (seed + threadIdx.x) & 0xff;
    for (int i = 0; i < 512; i++) { ge_p1p1 t; ge_dbl_p1p1(t, p.X, p.Y, p.Z); ge_dbl_fix(t); fe_mul(p.X, t.T, t.X); fe_mul(p.Y, t.Y, t.Z); fe_mul(p.Z, t.T, t.Z); }
    uint32_t r = 0; for (int i = 0; i < 10; i++) r ^= p.X.v[i] ^ p.Y.v[i] ^ p.Z.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class F> float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; void *buf; cudaMalloc(&buf, 64 << 20);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk);
    for (int tpb : {256, 512, 1024}) {
        int blocks = sms * (2048 / tpb) * 4; double n = (double)blocks * tpb;
        float ms = time_it([&] { k_imad_wide<<<blocks, tpb>>>((uint64_t *)buf, 12345, 678); });
        printf(", \"imad_wide_gops_t%d\": %.1f", tpb, n * ITERS * 8 / ms / 1e6);
        ms = time_it([&] { k_imad_lo<<<blocks, tpb>>>((uint32_t *)buf, 12345, 678); });
        printf(", \"imad_lo_gops_t%d\": %.1f", tpb, n * ITERS * 8 / ms / 1e6);
        ms = time_it([&] { k_dfma<<<blocks, tpb>>>((double *)buf, 1.0000001, 0.9999999); });
        printf(", \"dfma_gops_t%d\": %.1f", tpb, n * ITERS * 8 / ms / 1e6);
    }
    for (int tpb : {128, 256, 512}) {
        int blocks = sms * 8; double n = (double)blocks * tpb;
        float ms = time_it([&] { k_fe_mul<<<blocks, tpb>>>((uint32_t *)buf, 99); });
        printf(", \"fe_mul_gops_t%d\": %.2f", tpb, n * FE_ITERS * 2 / ms / 1e6);
        ms = time_it([&] { k_fe_sq<<<blocks, tpb>>>((uint32_t *)buf, 99); });
        printf(", \"fe_sq_gops_t%d\": %.2f", tpb, n * FE_ITERS * 2 / ms / 1e6);
        ms = time_it([&] { k_dbl<<<blocks, tpb>>>((uint32_t *)buf, 99); });
        printf(", \"dbl_gops_t%d\": %.3f", tpb, n * 512 / ms / 1e6);
    }
    printf("}\n");
    return 0;
}
