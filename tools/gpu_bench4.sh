#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for u in 3 2 4 0; do ROFL_UNFOLD=$u timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_u$u.json 2> gpurun_out/bench_u$u.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_u$u.json"))
    print("unfold=$u value=%.0f prove_ms=%.1f verify_ms=%.1f fold_ms=%.1f msm_ms=%.1f e2e=%.0f launches=%d" % (d["value"], d["prove_ms"], d["verify_ms"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["msm_ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
except Exception as ex: print("unfold=$u failed", ex, open("gpurun_out/bench_u$u.err").read()[-800:])
PY
done
ROFL_UNFOLD=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch3.log 2>&1
