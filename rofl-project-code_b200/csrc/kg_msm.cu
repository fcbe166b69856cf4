// kernel group MSM of kernels.cuh (one translation unit per group so the build runs in parallel)
#define KG_MSM 1
#include "kernels.cuh"
