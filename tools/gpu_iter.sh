#!/bin/bash
# one GPU iteration: parity tests, a bench line, and the ncu launch list of one prove+verify pass
set -x
mkdir -p gpurun_out
TAG=${1:-iter}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value=%.0f prove_ms=%.1f verify_ms=%.1f e2e=%.0f launches=%d" % (d["value"], d["prove_ms"], d["verify_ms"], d["e2e"]["value"], d["gpu_launches"]))
print(json.dumps(d["roofline"]))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
