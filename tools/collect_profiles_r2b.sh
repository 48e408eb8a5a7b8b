# Round-2 (second half) refresh of the C5 evidence after the half-storage K1: bench line with the CPU arm's parity block,
# launch list, full ncu capture of one C5 step (14 launches: K1 x4, prep, tql, replay, merge, tfactor, backwy, zgemm, nufft x2).
set -x
python bench.py > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_b_c5.err; tail -2 gpurun_out/r2_b_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hql_|tdc_|zgemm_|polar_" -c 13 -o /tmp/full_c5 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/full_c5.ncu-rep --page raw --csv > gpurun_out/r2_full_c5_raw.csv
python tools/show_bench.py gpurun_out/r2_bench_c5.json
