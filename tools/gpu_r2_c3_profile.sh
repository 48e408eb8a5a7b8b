# Source-level ncu capture of the d = 24 kernels (C3: ALC scan, integral mode), one launch of each kind.
ncu --set full --clock-control none --import-source on -k regex:"hql_|zgemm_" -c 6 -o /tmp/c3 python bench.py --workload c3 --n-orient 100 --steps 1 --warmup 0 --no-cpu > gpurun_out/c3_ncu.log 2>&1
ncu -i /tmp/c3.ncu-rep --page raw --csv > gpurun_out/c3_raw.csv
ncu -i /tmp/c3.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/c3_src.csv 2>/dev/null
ls -la gpurun_out/c3_*
