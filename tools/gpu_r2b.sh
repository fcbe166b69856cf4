#!/bin/bash
# r02 second round trip: correctness of the lane / priority-queue refactor, then a sweep of the scheduling knobs
set -x
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log; tail -3 gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q -k "absorb or weights or undecodable or clip or bytes_match or concurrent or full_size_round or config0" > gpurun_out/pytest_new_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_$TAG.log; tail -5 gpurun_out/pytest_new_$TAG.log
for cfg in "3 2" "3 1" "3 4" "3 0" "4 2" "2 2" "1 2"; do
  set -- $cfg
  ROFL_GROUPS=$1 ROFL_RT_PER=$2 BENCH_GROUPS=$1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g$1_p$2.json 2> gpurun_out/bench_${TAG}_g$1_p$2.err
  echo "groups=$1 rt_per=$2"; grep "resident per-step" gpurun_out/bench_${TAG}_g$1_p$2.err | cut -c1-200
done
