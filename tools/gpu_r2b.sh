# round 2, GPU call B: K4 split in two CTAs per matrix + separate T-factor kernel
set -x
python -m pytest tests/test_gpu_eigh.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2b_tests.log 2>&1; tail -5 gpurun_out/r2b_tests.log
python bench.py --no-cpu > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; tail -3 gpurun_out/r2b_bench_c5.err
python tools/show_bench.py gpurun_out/r2b_bench_c5.json
ncu --set full --clock-control none --import-source on -k regex:"hql_backwy|hql_tfactor" -c 2 -o /tmp/wy python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2b_ncu.log 2>&1
ncu -i /tmp/wy.ncu-rep --page raw --csv > gpurun_out/r2b_wy_raw.csv
python tools/ncu_full_summary.py gpurun_out/r2b_wy_raw.csv wy | grep -E "^##|duration|pipe active|warps active|stalls|DRAM"
