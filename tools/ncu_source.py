"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump: per-opcode instruction mix and
per-source-line stall samples (per-line figures are inclusive of inlined callees).  usage: ncu_source.py dump.csv [n_lines]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nshow = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
fpath = None
stalls = ["stall_barrier", "stall_wait", "stall_short_sb", "stall_math", "stall_not_selected", "stall_no_inst",
          "stall_branch_resolving", "stall_long_sb", "stall_selected", "stall_dispatch", "stall_mio", "stall_lg"]
bysrc = {}
opc = collections.Counter()
opsamp = collections.Counter()
tot = collections.Counter()
cur = None
seen = set()
for r in rows:
    if r and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        sidx = {s: hdr.index(s) for s in stalls}
        continue
    if not hdr or len(r) <= si:
        continue
    if r[2] == "-":  # a CUDA source line (aggregated over its SASS)
        cur = (fpath, r[0], r[1].strip()[:100])
        continue
    try:
        inst = int(r[ii] or 0)
        samp = int(r[si] or 0)
    except ValueError:
        continue
    first = r[2] not in seen  # inlined code is listed once per file of its inline stack
    seen.add(r[2])
    toks = r[3].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
    op = op.split(".")[0]
    if first:
        opc[op] += inst
        opsamp[op] += samp
        tot["inst"] += inst
        tot["samp"] += samp
    a = bysrc.setdefault(cur or ("?", "?", "?"), [0, 0, collections.Counter()])
    a[0] += inst
    a[1] += samp
    for s, j in sidx.items():
        v = int(r[j] or 0)
        a[2][s] += v
        if first:
            tot[s] += v
print("warp instructions %d, samples %d" % (tot["inst"], tot["samp"]))
print("stalls: " + "  ".join("%s %.1f%%" % (s[6:], 100 * tot[s] / tot["samp"]) for s in stalls))
print("--- by opcode")
for k, v in opc.most_common(24):
    print("%-10s %6.2f%% inst  %6.2f%% samples" % (k, 100 * v / tot["inst"], 100 * opsamp[k] / tot["samp"]))
print("--- by source line")
for k, a in sorted(bysrc.items(), key=lambda kv: -kv[1][1])[:nshow]:
    print("%5.1f%% smp %5.1f%% inst %s:%s %s | %s" % (100 * a[1] / tot["samp"], 100 * a[0] / tot["inst"], k[0], k[1], k[2][:80],
                                                    " ".join("%s=%.1f" % (s[6:], 100 * c / tot["samp"]) for s, c in a[2].most_common(3))))
