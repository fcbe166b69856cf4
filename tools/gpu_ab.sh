#!/bin/bash
# A/B bench runs under different env settings: tools/gpu_ab.sh TAG "ENV1=.. ENV2=.." "ENV.." ...
mkdir -p gpurun_out
TAG=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
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
