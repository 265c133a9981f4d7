#!/usr/bin/env python
"""Print selected metrics of every launch in an .ncu-rep (ncu -i X --page raw --csv).  Usage: ncu_raw.py X.ncu-rep [regex]"""
import csv, io, re, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
for r in rows[2:]:
    d = dict(zip(h, r)); u = dict(zip(h, units))
    if pat and not pat.search(d["Kernel Name"]):
        continue
    print("###", d["Kernel Name"][:90])
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k]} {u[k]} |")
    for k in h:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") or k.startswith("smsp__average_warp_latency_issue_stalled"):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v >= 0.3:
                print(f"| {k.replace('smsp__average_warps_issue_stalled_','stall ').replace('smsp__average_warp_latency_issue_stalled_','stall ').replace('_per_issue_active.ratio','').replace('.ratio','')} | {v:.2f} |")
