// IMAD.WIDE.U32 issue-rate probes whose operands change every iteration (ptxas cannot strength-reduce them).
//   wide_rot : c_k = lo(c_{k+1}) * y + c_k           (8 chains, distinct accumulators, one shared multiplier)
//   wide_rot2: c_k = lo(c_{k+1}) * lo(c_{k+2}) + c_k (both multiplicands vary)
//   wide_x   : carry-chained IMAD.WIDE.U32.X pairs as generated from mad.lo.cc / madc.hi.cc
//   lo_rot   : 32-bit IMAD with rotating operands
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void wide_rot(uint64_t *out, uint32_t y) {
    uint64_t c[8]; for (int k = 0; k < 8; k++) c[k] = threadIdx.x * 8 + k + 1;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = (uint64_t)(uint32_t)c[(k + 1) & 7] * y + c[k];
    }
    uint64_t r = 0; for (int k = 0; k < 8; k++) r ^= c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void wide_rot2(uint64_t *out, uint32_t y) {
    uint64_t c[8]; for (int k = 0; k < 8; k++) c[k] = threadIdx.x * 8 + k + y;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = (uint64_t)(uint32_t)c[(k + 1) & 7] * (uint32_t)c[(k + 2) & 7] + c[k];
    }
    uint64_t r = 0; for (int k = 0; k < 8; k++) r ^= c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void lo_rot(uint32_t *out, uint32_t y) {
    uint32_t c[8]; for (int k = 0; k < 8; k++) c[k] = threadIdx.x * 8 + k + y;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = c[(k + 1) & 7] * c[(k + 2) & 7] + c[k];
    }
    uint32_t r = 0; for (int k = 0; k < 8; k++) r ^= c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void wide_x(uint32_t *out, uint32_t y) {
    uint32_t a[8], t0 = threadIdx.x, t1 = 1, t2 = 2;
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 8 + k + y;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;" : "+r"(t0), "+r"(t1), "+r"(t2) : "r"(a[k]), "r"(a[(k + 3) & 7]));
        }
        a[i & 7] ^= t0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = t0 ^ t1 ^ t2;
}
template <class F> float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount; void *buf; cudaMalloc(&buf, 64 << 20);
    int tpb = 256, blocks = sms * 8 * 4; double n = (double)blocks * tpb;
    printf("{\"sms\": %d", sms);
    float ms = time_it([&] { wide_rot<<<blocks, tpb>>>((uint64_t *)buf, 12345); }); printf(", \"wide_rot_gops\": %.1f", n * ITERS * 8 / ms / 1e6);
    ms = time_it([&] { wide_rot2<<<blocks, tpb>>>((uint64_t *)buf, 12345); }); printf(", \"wide_rot2_gops\": %.1f", n * ITERS * 8 / ms / 1e6);
    ms = time_it([&] { lo_rot<<<blocks, tpb>>>((uint32_t *)buf, 12345); }); printf(", \"lo_rot_gops\": %.1f", n * ITERS * 8 / ms / 1e6);
    ms = time_it([&] { wide_x<<<blocks, tpb>>>((uint32_t *)buf, 12345); }); printf(", \"wide_x_gops\": %.1f", n * ITERS * 8 / ms / 1e6);
    printf("}\n"); return 0;
}
