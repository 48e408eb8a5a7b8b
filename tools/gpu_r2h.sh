# round 2, GPU call H: leaves through the batched QL kernels, faster secular loop
set -x
python -m pytest tests/test_gpu_eigh.py -m gpu -q -x > gpurun_out/r2h_tests_eigh.log 2>&1; tail -5 gpurun_out/r2h_tests_eigh.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2h_tests.log 2>&1; tail -5 gpurun_out/r2h_tests.log
python bench.py --no-cpu > gpurun_out/r2h_bench_c5.json 2> gpurun_out/r2h_bench_c5.err; tail -3 gpurun_out/r2h_bench_c5.err
python tools/show_bench.py gpurun_out/r2h_bench_c5.json
python bench.py --no-cpu --n-orient 2500 > gpurun_out/r2h_bench_c5_2500.json 2> gpurun_out/r2h_2500.err
python tools/show_bench.py gpurun_out/r2h_bench_c5_2500.json
ncu --set full --clock-control none --import-source on -k regex:"tdc_merge|hql_tfactor|hql_backwy" -c 3 -o /tmp/tdc python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2h_ncu.log 2>&1
ncu -i /tmp/tdc.ncu-rep --page raw --csv > gpurun_out/r2h_raw.csv
ncu -i /tmp/tdc.ncu-rep --page source --csv > gpurun_out/r2h_src.csv
python tools/ncu_full_summary.py gpurun_out/r2h_raw.csv tdc | grep -E "^##|duration|pipe active|warps active|stalls|LSU|regs|local"
