#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "config3 or config4 or bytes_match or full_size_round or oracle_proofs" --durations=5 > gpurun_out/pytest_new_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_r2g.log; tail -12 gpurun_out/pytest_new_r2g.log
ROFL_ALLOC_TRACE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; grep "per-step" gpurun_out/bench_r2g.err | cut -c1-260; grep -c "rofl alloc" gpurun_out/bench_r2g.err
ROFL_TIMELINE=1 ROFL_GROUPS=1 timeout 600 python tools/timeline_cfg3.py gpurun_out/tl_cfg3_g1_b.txt 8 2>&1 | tail -3
ROFL_GROUPS=3 timeout 300 python tools/timeline_cfg3.py /dev/null 8 2>&1 | tail -2
ROFL_GROUPS=1 timeout 300 python tools/timeline_cfg3.py /dev/null 8 2>&1 | tail -2
