#!/bin/bash
# like gpu_ab.sh without the pytest pass, plus an ncu launch list for the LAST env setting
mkdir -p gpurun_out
TAG=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_$i.json"))
r=d["roofline"]
print("[$envs] value=%.0f prove_ms=%.2f verify_ms=%.2f e2e=%.0f launches=%d fold=%.1f msm=%.1f" % (d["value"], d["prove_ms"], d["verify_ms"], d["e2e"]["value"], d["gpu_launches"], r.get("kernel_ms_per_step") or 0, r.get("msm_ms_per_step") or 0))
PY
done
env $envs timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
