// kernel group SQUARE of kernels.cuh (one translation unit per group so the build runs in parallel)
#define KG_SQUARE 1
#include "kernels.cuh"
