#!/bin/bash
# round-end rehearsal: what the driver runs (gpu tests, smoke, bench both arms) + the profile artefacts
set -x
mkdir -p gpurun_out
TAG=${1:-final}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cut -c1-300 gpurun_out/bench_ref_$TAG.json
ROFL_GROUPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
ROFL_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_rt_msm$ -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_k_rt_msm python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_${TAG}_rt.log 2>&1
