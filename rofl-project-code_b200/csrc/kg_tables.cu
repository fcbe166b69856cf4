// kernel group TABLES of kernels.cuh (one translation unit per group so the build runs in parallel)
#define KG_TABLES 1
#include "kernels.cuh"
