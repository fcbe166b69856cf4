// Field / point arithmetic throughput against resident warps per SM: how much of the IMAD.WIDE pipe a chain of dependent
// multiplications (fe_mul), two interleaved chains, and a chain of mixed additions (ge_madd, 7 M) reach at 4..32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rofl-project-code_b200/csrc -I include tools/microbench3.cu -o tools/microbench3
#include <cstdio>
#include <cuda_runtime.h>
#define KG_TABLES 1
#include "kernels.cuh"
void rt_count_launch(const char *) {}
void *rt_prof_begin(int, cudaStream_t) { return nullptr; }
void rt_prof_end(int, void *, cudaStream_t) {}
#define ITERS 2048
__global__ void k_mul1(uint32_t *out) {
    fe x, y; for (int i = 0; i < 8; i++) { x.v[i] = threadIdx.x * 8 + i + 1; y.v[i] = blockIdx.x + i + 3; }
    for (int i = 0; i < ITERS; i++) fe_mul(x, x, y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ x.v[7];
}
__global__ void k_mul2(uint32_t *out) {
    fe x, y, z; for (int i = 0; i < 8; i++) { x.v[i] = threadIdx.x * 8 + i + 1; y.v[i] = blockIdx.x + i + 3; z.v[i] = x.v[i] ^ 0x55; }
    for (int i = 0; i < ITERS / 2; i++) { fe_mul(x, x, y); fe_mul(z, z, y); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ z.v[7];
}
__global__ void k_madd(uint32_t *out) {
    ge_p3 B; ge_base(B); ge_p3 acc = B; fe zi; fe_invert(zi, B.Z); ge_niels nb; ge_p3_to_niels(nb, B, zi);
    nb.xy2d.v[0] ^= threadIdx.x;                    // not a valid point any more: only the instruction stream matters
    for (int i = 0; i < ITERS / 8; i++) ge_madd(acc, acc, nb);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.X.v[0] ^ acc.T.v[7];
}
template <class F> float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount; void *buf; cudaMalloc(&buf, 64 << 20);
    printf("warps/SM : fe_mul chain | two chains | ge_madd chain   (T IMAD.WIDE-equivalents/s at 72 per multiplication)\n");
    for (int bps : {1, 2, 3, 4, 6, 8}) {
        int blocks = sms * bps; double n = (double)blocks * 128;
        float a = time_it([&] { k_mul1<<<blocks, 128>>>((uint32_t *)buf); });
        float b = time_it([&] { k_mul2<<<blocks, 128>>>((uint32_t *)buf); });
        float c = time_it([&] { k_madd<<<blocks, 128>>>((uint32_t *)buf); });
        printf("%2d : %.2f | %.2f | %.2f\n", bps * 4, n * ITERS * 72 / a / 1e9, n * ITERS * 72 / b / 1e9, n * (ITERS / 8) * 7 * 72 / c / 1e9);
    }
    return 0;
}
