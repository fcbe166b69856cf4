#!/bin/bash
# GPU round trip: parity tests, bench line, ncu launch list, ncu full capture of the dominant kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ipp_fold_points -c 1 -o gpurun_out/prof_fold python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_fold.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fe_mul -c 1 -o gpurun_out/prof_femul ./tools/microbench > gpurun_out/ncu_femul.log 2>&1
ls -la gpurun_out
