#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "bytes_match or full_size_round or config4_server_batch_path or config3" > gpurun_out/pytest_new_r2l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_r2l.log; tail -3 gpurun_out/pytest_new_r2l.log
ROFL_TRACE=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err
grep "per-step" gpurun_out/bench_r2l.err | cut -c1-300; grep "rofl trace" gpurun_out/bench_r2l.err | grep "v_tr" | tail -2
