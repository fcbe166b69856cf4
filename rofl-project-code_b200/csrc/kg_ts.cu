// kernel group TS (device-side Merlin transcripts, ts_kernels.cuh) of kernels.cuh
#define KG_TS 1
#include "kernels.cuh"
