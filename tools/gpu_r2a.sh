# round 2, GPU call A: tests, C5 bench with the compact-WY K4 (default) and the level-2 K4, ncu of the new kernel
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -q > gpurun_out/r2a_tests.log 2>&1; tail -15 gpurun_out/r2a_tests.log
python bench.py > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err; tail -3 gpurun_out/r2a_bench_c5.err
python bench.py --no-cpu --steps 3 --option back_wy=0 > gpurun_out/r2a_bench_c5_nowy.json 2> gpurun_out/r2a_nowy.err
ncu --set full --clock-control none --import-source on -k regex:hql_backwy -c 1 -o /tmp/wy python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2a_ncu.log 2>&1
ncu -i /tmp/wy.ncu-rep --page raw --csv > gpurun_out/r2a_wy_raw.csv
ncu -i /tmp/wy.ncu-rep --page source --csv > gpurun_out/r2a_wy_src.csv
python tools/show_bench.py gpurun_out/r2a_bench_c5.json; python tools/show_bench.py gpurun_out/r2a_bench_c5_nowy.json
