# Final C5 lines after the last kernel changes: full line with the CPU arm, 2 500-orientation shard, launch list.
set -x
python bench.py > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_b_c5.err; tail -2 gpurun_out/r2_b_c5.err
python bench.py --n-orient 2500 --no-cpu > gpurun_out/r2_bench_c5_2500.json 2> gpurun_out/r2_b_c5_2500.err
python bench.py --general > gpurun_out/r2_bench_c5_general.json 2> gpurun_out/r2_b_c5g.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hql_|tdc_|zgemm_|polar_" -c 13 -o /tmp/full_c5 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/full_c5.ncu-rep --page raw --csv > gpurun_out/r2_full_c5_raw.csv
for f in gpurun_out/r2_bench_c5.json gpurun_out/r2_bench_c5_2500.json gpurun_out/r2_bench_c5_general.json; do python tools/show_bench.py $f; done
