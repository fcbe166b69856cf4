#!/bin/bash
# ncu --set full captures of named kernels of one prove+verify pass: tools/gpu_prof.sh TAG "kernel:skip" ...
set -x
mkdir -p gpurun_out
TAG=$1; shift
for spec in "$@"; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^${k}$ -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_${k}_${skip} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_${TAG}_${k}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
