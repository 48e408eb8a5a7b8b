# round 2, GPU call E: one-barrier K1, faster secular loop
set -x
python -m pytest tests/test_gpu_eigh.py -m gpu -q -x > gpurun_out/r2e_tests_eigh.log 2>&1; tail -5 gpurun_out/r2e_tests_eigh.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2e_tests.log 2>&1; tail -5 gpurun_out/r2e_tests.log
python bench.py --no-cpu > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err; tail -3 gpurun_out/r2e_bench_c5.err
python tools/show_bench.py gpurun_out/r2e_bench_c5.json
python bench.py --no-cpu --option tridiag_one=0 --steps 3 > gpurun_out/r2e_bench_c5_two.json 2> gpurun_out/r2e_two.err
python tools/show_bench.py gpurun_out/r2e_bench_c5_two.json
python bench.py --no-cpu --option tdc=0 --steps 3 > gpurun_out/r2e_bench_c5_notdc.json 2> gpurun_out/r2e_notdc.err
python tools/show_bench.py gpurun_out/r2e_bench_c5_notdc.json
ncu --set full --clock-control none --import-source on -k regex:"tdc_|hql_tridiag_rw1_kernel" -c 5 -o /tmp/tdc python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2e_ncu.log 2>&1
ncu -i /tmp/tdc.ncu-rep --page raw --csv > gpurun_out/r2e_raw.csv
ncu -i /tmp/tdc.ncu-rep --page source --csv > gpurun_out/r2e_src.csv
python tools/ncu_full_summary.py gpurun_out/r2e_raw.csv tdc | grep -E "^##|duration|pipe active|warps active|stalls|LSU|regs|local"
