"""Summarise `ncu --page raw --csv` output of an `ncu --set full` capture: one table per kernel launch.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; ncu_full_summary.py raw.csv "title" [out.json n_cfg] > out.md

With `out.json n_cfg` the per-launch numbers bench.py needs (duration, DRAM bytes, pipe utilisation)
are also written as JSON; n_cfg = configurations per launch group in the captured command.  bench.py
reads profiles/r2_ncu_<workload>.json for `roofline.traffic`."""
import csv
import json
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

if len(sys.argv) > 4:
    def num(d, k):
        try:
            return float(d.get(k, "") or 0)
        except ValueError:
            return 0.0

    def to_bytes(d, k):
        v, u = num(d, k), unit.get(k, "byte").lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)

    def to_ns(d, k):
        v, u = num(d, k), unit.get(k, "ns").lower()
        return v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9, "second": 1e9}.get(u, 1.0)

    ks = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        ks.append({
            "name": d["Kernel Name"].split("(")[0].replace("void ", ""),
            "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size"),
            "duration_ns": to_ns(d, "gpu__time_duration.sum"),
            "dram_read_bytes": to_bytes(d, "dram__bytes_read.sum"),
            "dram_write_bytes": to_bytes(d, "dram__bytes_write.sum"),
            "regs": num(d, "launch__registers_per_thread"),
            "fp64_pipe_pct": num(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "dmma_pipe_pct": num(d, "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
            "lsu_pct": num(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "local_load_sectors": num(d, "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
        })
    json.dump({"title": title, "configurations": int(sys.argv[4]), "kernels": ks}, open(sys.argv[3], "w"), indent=1)
