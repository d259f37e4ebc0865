#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [KERNEL_SUBSTRING] > profiles/<name>.md  — headline raw metrics of one kernel of an
`ncu --set full` capture as a markdown table (the per-region source attribution comes from tools/prof_lines.py)."""
import csv, os, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name")
vals = [r for r in rows[2:] if want in r[kn]][0]
print("kernel: `%s`\n" % vals[kn])
keep = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
print("| metric | value | unit |\n|---|---|---|")
for i, h in enumerate(hdr):
    if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print("| %s | %s | %s |" % (h, vals[i], units[i]))
