// kernel group SCALAR of kernels.cuh (one translation unit per group so the build runs in parallel)
#define KG_SCALAR 1
#include "kernels.cuh"
