set -x
python bench.py > gpurun_out/r1_bench_c5.json 2> gpurun_out/b5.err
python bench.py --workload c2 > gpurun_out/r1_bench_c2.json 2> gpurun_out/b2.err
python bench.py --workload c3 > gpurun_out/r1_bench_c3.json 2> gpurun_out/b3.err
python bench.py --workload c4 > gpurun_out/r1_bench_c4.json 2> gpurun_out/b4.err
python bench.py --general > gpurun_out/r1_bench_c5_general.json 2> gpurun_out/b5g.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:hql_|zgemm_|polar_" -c 14 -o gpurun_out/full_c5 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/b_ncu2.log 2>&1
ncu -i gpurun_out/full_c5.ncu-rep --page raw --csv > gpurun_out/full_c5_raw.csv
for f in gpurun_out/r1_bench_c*.json; do python tools/show_bench.py $f; done
