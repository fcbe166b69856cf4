#!/bin/bash
# r02: full gpu suite, bench (with configs block), reference arm, launch list, ncu captures of k_rt_msm / k_msm
set -x
mkdir -p gpurun_out
TAG=${1:-r2e}
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log; tail -20 gpurun_out/pytest_$TAG.log
timeout 1500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -25 gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cut -c1-300 gpurun_out/bench_ref_$TAG.json
ROFL_GROUPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_launch_$TAG.log 2>&1
ROFL_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_rt_msm$ -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_k_rt_msm python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra > gpurun_out/ncu_${TAG}_rt.log 2>&1
# configs[3] (no generator tables): launch list + ncu captures of the wide-window bucket MSM and the catch-up fold, on 2 chunks of 2^18 values
ROFL_GROUPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg3_$TAG.csv python tools/timeline_cfg3.py /dev/null 2 > gpurun_out/ncu_launch_cfg3_$TAG.log 2>&1
ROFL_GROUPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_bm_accum|k_catchup_naf)$' -s 2 -c 2 -f -o gpurun_out/prof_${TAG}_cfg3 python tools/timeline_cfg3.py /dev/null 2 > gpurun_out/ncu_${TAG}_cfg3.log 2>&1
