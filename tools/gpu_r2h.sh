#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "config3 or bytes_match or oracle_proofs or config4_server_batch_path" > gpurun_out/pytest_new_r2h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_r2h.log; tail -4 gpurun_out/pytest_new_r2h.log
ROFL_TIMELINE=1 ROFL_GROUPS=1 timeout 600 python tools/timeline_cfg3.py gpurun_out/tl_cfg3_g1_c.txt 8 2>&1 | tail -3
ROFL_GROUPS=3 timeout 300 python tools/timeline_cfg3.py /dev/null 8 2>&1 | tail -2
ROFL_GROUPS=1 timeout 300 python tools/timeline_cfg3.py /dev/null 8 2>&1 | tail -2
ROFL_GROUPS=2 timeout 300 python tools/timeline_cfg3.py /dev/null 8 2>&1 | tail -2
