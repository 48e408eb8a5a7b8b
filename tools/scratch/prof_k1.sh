ncu --set full --clock-control none --import-source on -k regex:"hql_tridiag_hs" -c 1 -o /tmp/k1 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/hs_ncu_full.log 2>&1
ncu -i /tmp/k1.ncu-rep --page raw --csv > gpurun_out/hs_k1_raw.csv
ncu -i /tmp/k1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/hs_k1_src.csv 2>/dev/null
ncu -i /tmp/k1.ncu-rep --page source --csv --print-source sass > gpurun_out/hs_k1_sass.csv 2>/dev/null
python tools/ncu_full_summary.py gpurun_out/hs_k1_raw.csv "K1 phases" > gpurun_out/hs_k1_summary.md
