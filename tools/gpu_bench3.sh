#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for g in 1 2; do ROFL_GROUPS=$g timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_g$g.json"))
print("groups=$g value=%.0f prove_ms=%.1f verify_ms=%.1f fold_ms=%.1f msm_ms=%.1f e2e=%.0f launches=%d" % (d["value"], d["prove_ms"], d["verify_ms"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["msm_ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
PY
done
ROFL_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm -c 1 -o gpurun_out/prof_msm python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_msm.log 2>&1
ROFL_GROUPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1
