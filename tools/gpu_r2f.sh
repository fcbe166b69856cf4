#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "absorb or harness or bindings32 or bytes_match" > gpurun_out/pytest_new_r2f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_r2f.log; tail -5 gpurun_out/pytest_new_r2f.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; grep "per-step" gpurun_out/bench_r2f.err | cut -c1-220
ROFL_TIMELINE=1 ROFL_GROUPS=1 timeout 600 python tools/timeline_cfg3.py gpurun_out/tl_cfg3_g1.txt 8 2>&1 | tail -3
ROFL_GROUPS=1 timeout 300 python tools/timeline_cfg3.py /dev/null 8 2>&1 | tail -2
ROFL_GROUPS=3 timeout 300 python tools/timeline_cfg3.py /dev/null 8 2>&1 | tail -2
ROFL_TIMELINE=1 ROFL_GROUPS=1 timeout 300 python tools/timeline_run.py gpurun_out/tl_g1_b.txt
