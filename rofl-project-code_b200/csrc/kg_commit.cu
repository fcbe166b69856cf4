// kernel group COMMIT of kernels.cuh (one translation unit per group so the build runs in parallel)
#define KG_COMMIT 1
#include "kernels.cuh"
