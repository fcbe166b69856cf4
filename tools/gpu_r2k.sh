#!/bin/bash
set -x
mkdir -p gpurun_out
ROFL_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2k_extra.json 2> gpurun_out/bench_r2k_extra.err
ROFL_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/bench_r2k_noextra.json 2> gpurun_out/bench_r2k_noextra.err
grep "per-step" gpurun_out/bench_r2k_extra.err gpurun_out/bench_r2k_noextra.err | cut -c1-200
