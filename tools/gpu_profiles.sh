#!/bin/bash
# round profile artefacts: ncu launch list of one prove+verify pass and a full capture of the dominant kernel
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
ROFL_GROUPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
ROFL_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_rt_msm$ -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_k_rt_msm python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_${TAG}_rt.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -2 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
cat gpurun_out/bench_ref_$TAG.json | cut -c1-600
