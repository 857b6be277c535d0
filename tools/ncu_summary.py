#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md argues from."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued", "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__sass_inst_executed_op_local", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg ", "launch__waves", "dram__bytes_read.sum.per_second", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_srcunit_tex.sum"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:80])
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(k.strip()) if k.endswith(" ") else k in h for k in KEYS) and "max" not in h and "min" not in h and ".sum.p" not in h.replace(".sum.per_second", ""):
            print(f"  {h} [{u}] = {v}")
