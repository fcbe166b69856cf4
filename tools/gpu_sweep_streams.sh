#!/bin/bash
# "groups side maxconn"
mkdir -p gpurun_out
TAG=${1:-sweep}; shift
STEPS=${STEPS:-16}
: > gpurun_out/sweep_$TAG.txt
i=0
for cfg in "$@"; do
  set -- $cfg; i=$((i+1))
  if [ "$3" != "-" ]; then export CUDA_DEVICE_MAX_CONNECTIONS=$3; else unset CUDA_DEVICE_MAX_CONNECTIONS; fi
  ROFL_GROUPS=$1 ROFL_SIDE=$2 BENCH_GROUPS=$1 timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  echo "groups=$1 side=$2 maxconn=$3" >> gpurun_out/sweep_$TAG.txt; grep "resident per-step" gpurun_out/bench_${TAG}_$i.err | cut -c1-420 >> gpurun_out/sweep_$TAG.txt
done
cat gpurun_out/sweep_$TAG.txt
