#!/bin/bash
# r02 third round trip: ordered bulk priorities -- sweep of groups / split / rt_per
set -x
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 600 python -m pytest tests -m gpu -x -q -k "bytes_match or concurrent or full_size_round" > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_$TAG.log; tail -3 gpurun_out/pytest_new_$TAG.log
i=0
for cfg in "3 2 1,1,1" "3 2 3,2,1" "3 2 4,3,2" "2 2 1,1,1" "2 2 2,1,1" "4 2 4,3,2,1" "3 1 3,2,1" "3 4 3,2,1" "4 2 3,3,2,1"; do
  set -- $cfg; i=$((i+1))
  ROFL_GROUPS=$1 ROFL_RT_PER=$2 ROFL_SPLIT=$3 BENCH_GROUPS=$1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  echo "groups=$1 rt_per=$2 split=$3" | tee -a gpurun_out/sweep_$TAG.txt; grep "resident per-step" gpurun_out/bench_${TAG}_$i.err | cut -c1-200 | tee -a gpurun_out/sweep_$TAG.txt
done
