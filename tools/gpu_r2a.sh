#!/bin/bash
# r02 first GPU round trip: smoke, the new device-transcript tests first, bench, full gpu suite, launch list
set -x
mkdir -p gpurun_out
TAG=${1:-r2a}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log; tail -3 gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q -k "absorb or weights or undecodable or clip or bytes_match or field or commit_conv" > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_$TAG.log; tail -5 gpurun_out/pytest_new_$TAG.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -12 gpurun_out/bench_$TAG.err; cut -c1-600 gpurun_out/bench_$TAG.json
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log; tail -25 gpurun_out/pytest_$TAG.log
ROFL_GROUPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
