#!/usr/bin/env python3
"""Key metrics of an .ncu-rep (ncu --set full): duration, occupancy, pipe utilisation, stall reasons, memory traffic."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy", "sm__pipe_fmaheavy_cycles_active", "sm__inst_executed_pipe_alu", "sm__pipe_alu_cycles_active", "smsp__inst_executed.sum ",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "lts__t_bytes.sum ", "l1tex__t_bytes.sum ", "smsp__average_warps_issue_stalled", "sass__inst_executed_local", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct", "smsp__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_fma", "achieved_occupancy", "sm__cycles_active.avg "]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", rep, r[h.index("Kernel Name")])
        for i, n in enumerate(h):
            if any(k.strip() in n for k in KEYS) and r[i] not in ("", "0"):
                try:
                    v = float(r[i])
                    if "stalled" in n and v < 0.15: continue
                except ValueError:
                    pass
                print(f"  {n} = {r[i]} {units[i]}")
