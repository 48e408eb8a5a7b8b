# Collect the bench lines and ncu captures that profiles/r2_* are built from (run on the GPU box via gpurun).
# .ncu-rep files stay in /tmp on the box (gpurun_out/ is capped at 64 MiB); only the CSV pages come back.
set -x
for w in c5 c2 c3 c4; do python bench.py --workload $w > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_b_$w.err; tail -2 gpurun_out/r2_b_$w.err; done
python bench.py --general > gpurun_out/r2_bench_c5_general.json 2> gpurun_out/r2_b_c5g.err
python bench.py --workload c2 --general > gpurun_out/r2_bench_c2_general.json 2> gpurun_out/r2_b_c2g.err
python bench.py --n-orient 2500 --no-cpu > gpurun_out/r2_bench_c5_2500.json 2> gpurun_out/r2_b_c5_2500.err
# launch list of one C5 step (cold-cache, serialised: compare SHARES with the event timers)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_launch.log 2>&1
# full capture of every kernel of one C5 step (12 launches: K1 x3, prep, tql, replay, merge, tfactor, backwy, zgemm, nufft x2)
ncu --set full --clock-control none --import-source on -k regex:"hql_|tdc_|zgemm_|polar_" -c 12 -o /tmp/full_c5 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/full_c5.ncu-rep --page raw --csv > gpurun_out/r2_full_c5_raw.csv
# the FP64 peak micro-benchmarks behind the roofline denominator
ncu --set full --clock-control none -k regex:"peak_" -c 2 -o /tmp/peak python -c "
from muspinsim_b200 import _lib
print(_lib.fp64_peak(0,0), _lib.fp64_peak(0,1))" > gpurun_out/r2_ncu_peak.log 2>&1
ncu -i /tmp/peak.ncu-rep --page raw --csv > gpurun_out/r2_peak_raw.csv
# C4 (Lindbladian) kernels
ncu --set full --clock-control none -k regex:"lind_" -c 4 -o /tmp/full_c4 python bench.py --workload c4 --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_c4.log 2>&1
ncu -i /tmp/full_c4.ncu-rep --page raw --csv > gpurun_out/r2_full_c4_raw.csv
for f in gpurun_out/r2_bench_c*.json; do python tools/show_bench.py $f; done
