// kernel group FOLD of kernels.cuh (one translation unit per group so the build runs in parallel)
#define KG_FOLD 1
#include "kernels.cuh"
