#!/bin/bash
# first GPU round trip: smoke, integer-pipe microbenchmarks, GPU parity tests
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 ./tools/microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/smoke.log; cat gpurun_out/microbench.json; tail -15 gpurun_out/pytest_gpu.log
