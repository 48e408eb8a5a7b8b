# Collect the bench lines and ncu captures that profiles/ is built from (run on the GPU box via gpurun).
# .ncu-rep files stay in /tmp on the box (gpurun_out/ is capped at 64 MiB); only the CSV pages come back.
set -x
WL="${1:-c5 c2 c3 c4}"
for w in $WL; do python bench.py --workload $w > gpurun_out/r1_bench_$w.json 2> gpurun_out/b_$w.err; done
python bench.py --general > gpurun_out/r1_bench_c5_general.json 2> gpurun_out/b5g.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:hql_|zgemm_|polar_" -c 9 -o /tmp/full_c5 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/b_ncu2.log 2>&1
ncu -i /tmp/full_c5.ncu-rep --page raw --csv > gpurun_out/full_c5_raw.csv
ncu --set full --clock-control none -k "regex:rho0_|zgemm_" -c 5 -o /tmp/full_c5g python bench.py --general --steps 1 --warmup 0 --no-cpu > gpurun_out/b_ncu3.log 2>&1
ncu -i /tmp/full_c5g.ncu-rep --page raw --csv > gpurun_out/full_c5g_raw.csv
for f in gpurun_out/r1_bench_c*.json; do python tools/show_bench.py $f; done
