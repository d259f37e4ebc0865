#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep > profiles/<name>.md  — headline raw metrics of one `ncu --set full` capture as a
markdown table (the per-region source attribution comes from tools/prof_report.py)."""
import csv, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
print("| metric | value | unit |\n|---|---|---|")
for i, h in enumerate(hdr):
    if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print("| %s | %s | %s |" % (h, vals[i], units[i]))
