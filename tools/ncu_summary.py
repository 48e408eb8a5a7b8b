"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(per kernel: launches, total ms, share).  usage: ncu_summary.py launches.csv [title] > out.md"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if k.startswith("musim::") and "peak_" not in k}
    tot = sum(v[1] for v in ours.values())
    print("# %s\n" % title)
    print("ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`): cold-cache, serialised;")
    print("compare SHARES, not absolutes.  Only this library's kernels (peak micro-benchmarks excluded).\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, t, 100 * t / tot))
    print("\nother kernels in the capture: " + ", ".join("%s (%d)" % (k[:40], v[0]) for k, v in agg.items() if k not in ours))


main()
