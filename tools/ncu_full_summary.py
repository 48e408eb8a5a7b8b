"""Summarise `ncu --page raw --csv` output of an `ncu --set full` capture: one table per kernel launch.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; ncu_full_summary.py raw.csv "title" > out.md"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
hdr, units = rows[0], rows[1]
unit = dict(zip(hdr, units))
KEYS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid x block", None),
    ("regs/thread", "launch__registers_per_thread"),
    ("CTAs/SM limit (regs, smem)", None),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("FP64 pipe active %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("DMMA pipe active %", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("LSU wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("local-memory load sectors", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
]
print("# %s\n" % title)
stall_cols = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    print("## `%s`\n" % name)
    print("| metric | value |\n|---|---|")
    for label, key in KEYS:
        if label == "grid x block":
            print("| grid x block | %s x %s |" % (d.get("launch__grid_size"), d.get("launch__block_size")))
        elif key is None:
            print("| %s | %s, %s |" % (label, d.get("launch__occupancy_limit_registers"), d.get("launch__occupancy_limit_shared_mem")))
        elif key in d:
            print("| %s | %s %s |" % (label, d[key], unit.get(key, "")))
    st = sorted(((float(d[h] or 0), h.split("issue_stalled_")[1].split("_per")[0]) for h in stall_cols), reverse=True)[:5]
    print("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (b, a) for a, b in st))
    print()
