# Final round-2 evidence after the half-grid NUFFT and the third-order secular step (one B200): bench lines of every
# workload with the CPU arm's parity block, launch list and full ncu capture of one C5 step.
set -x
for w in c5 c2 c3 c4; do python bench.py --workload $w > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_b_$w.err; tail -2 gpurun_out/r2_b_$w.err; done
python bench.py --general > gpurun_out/r2_bench_c5_general.json 2> gpurun_out/r2_b_c5g.err
python bench.py --workload c2 --general > gpurun_out/r2_bench_c2_general.json 2> gpurun_out/r2_b_c2g.err
python bench.py --n-orient 2500 --no-cpu > gpurun_out/r2_bench_c5_2500.json 2> gpurun_out/r2_b_c5_2500.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hql_|tdc_|zgemm_|polar_" -c 13 -o /tmp/full_c5 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/full_c5.ncu-rep --page raw --csv > gpurun_out/r2_full_c5_raw.csv
for f in gpurun_out/r2_bench_c*.json; do python tools/show_bench.py $f; done
