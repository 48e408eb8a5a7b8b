# Source-level ncu capture of the first K1 phase (half-storage DMMA tridiagonalisation, live block 96 -> 72) at C5.
ncu --set full --clock-control none --import-source on -k regex:"hql_tridiag_hs" -c 1 -o /tmp/k1 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/k1_ncu.log 2>&1
ncu -i /tmp/k1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/k1_src.csv 2>/dev/null
ls -la gpurun_out/k1_src.csv
