"""Compact view of a bench.py JSON line (stdin or file args)."""
import json
import sys

srcs = [open(a) for a in sys.argv[1:]] or [sys.stdin]
for f in srcs:
    for line in f:
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print("%s: value %.0f %s  %.2f ms/step  e2e %.0f  launches %s" % (
            d.get("config", {}).get("workload"), d["value"], d["unit"], d["ms_per_step"],
            d.get("e2e", {}).get("value", 0), d.get("gpu_launches")))
        ph = d.get("kernel_ms_per_step") or {}
        print("   " + "  ".join("%s %.2f" % (k, v) for k, v in ph.items() if v))
        r = d.get("roofline") or {}
        if r:
            print("   roofline: %s %.2f/%.1f %s frac %.3f | path frac %.3f" % (
                r.get("kernel"), r.get("achieved", 0), r.get("peak", 0), r.get("unit"), r.get("frac") or 0,
                (r.get("path") or {}).get("frac") or 0))
        if d.get("cpu_baseline"):
            print("   cpu:", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"), d["cpu_baseline"].get("kind"))
        print("   clocks:", d.get("clocks"))
