#!/bin/bash
# step-time noise A/B: tools/noise_ab.sh ROUNDS "ENV_A" "ENV_B" ... ; prints median / mean / slow-step count per run
R=$1; shift
for r in $(seq 1 $R); do for envs in "$@"; do
  env $envs BENCH_NO_CLOCKS=1 timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep -E "^(resident|e2e) per-step" | python -c "
import sys,re,statistics as st
tot=[]
for l in sys.stdin:
    v=[float(x) for x in re.findall(r'\((\d+\.\d+),', l)]; tot+=v
print('[$envs] median %.1f mean %.1f slow(>1.2x) %d/%d' % (st.median(tot), st.mean(tot), sum(1 for x in tot if x>1.2*st.median(tot)), len(tot)))
"; done; done
