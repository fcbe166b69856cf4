#!/bin/bash
set -x
mkdir -p gpurun_out
export ROFL_TIMELINE=1
ROFL_GROUPS=1 timeout 300 python tools/timeline_run.py gpurun_out/tl_g1.txt
ROFL_GROUPS=3 ROFL_PRIO_ORDERED=0 timeout 300 python tools/timeline_run.py gpurun_out/tl_g3_equal.txt
ROFL_GROUPS=3 ROFL_PRIO_ORDERED=1 timeout 300 python tools/timeline_run.py gpurun_out/tl_g3_ordered.txt
ROFL_GROUPS=3 ROFL_PRIO_ORDERED=0 ROFL_PRIO_FLAT=1 timeout 300 python tools/timeline_run.py gpurun_out/tl_g3_flat.txt
ROFL_GROUPS=2 ROFL_PRIO_ORDERED=1 ROFL_SPLIT=2,1 timeout 300 python tools/timeline_run.py gpurun_out/tl_g2_ordered.txt
