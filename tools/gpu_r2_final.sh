#!/bin/bash
# Round-2 late session: quick A/B of the bench lines after a kernel change (no CPU arm).
set -x
T=${1:-r2m}
for w in c5 c2; do python bench.py --workload $w --no-cpu > gpurun_out/${T}_$w.json 2>/dev/null; python tools/show_bench.py gpurun_out/${T}_$w.json | head -2; done
python bench.py --workload c5 --n-orient 2500 --no-cpu > gpurun_out/${T}_c5_2500.json 2>/dev/null; python tools/show_bench.py gpurun_out/${T}_c5_2500.json | head -2
python bench.py --workload c2 --general --no-cpu > gpurun_out/${T}_c2g.json 2>/dev/null; python tools/show_bench.py gpurun_out/${T}_c2g.json | head -2
if [ -n "$2" ]; then python bench.py --workload c3 --no-cpu > gpurun_out/${T}_c3.json 2>/dev/null; python tools/show_bench.py gpurun_out/${T}_c3.json | head -2; fi
