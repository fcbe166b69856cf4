// kernel group BSGS of kernels.cuh (one translation unit per group so the build runs in parallel)
#define KG_BSGS 1
#include "kernels.cuh"
