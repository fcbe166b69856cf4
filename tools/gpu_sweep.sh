#!/bin/bash
# sweep of the chunk-group scheduling knobs on configs[1] (bench.py --no-extra): "groups rt_per split ordered flat"
mkdir -p gpurun_out
TAG=${1:-sweep}; shift
STEPS=${STEPS:-6}
: > gpurun_out/sweep_$TAG.txt
i=0
for cfg in "$@"; do
  set -- $cfg; i=$((i+1))
  ROFL_GROUPS=$1 ROFL_RT_PER=$2 ROFL_SPLIT=$3 ROFL_PRIO_ORDERED=$4 ROFL_PRIO_FLAT=${5:-0} BENCH_GROUPS=$1 timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  echo "groups=$1 rt_per=$2 split=$3 ordered=$4 flat=${5:-0}" >> gpurun_out/sweep_$TAG.txt; grep "resident per-step" gpurun_out/bench_${TAG}_$i.err | cut -c1-420 >> gpurun_out/sweep_$TAG.txt
done
cat gpurun_out/sweep_$TAG.txt
