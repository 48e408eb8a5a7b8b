# Source-level ncu capture of K4 (T factors + compact-WY back-transformation) and K5 (rotation GEMM) at C5.
ncu --set full --clock-control none --import-source on -k regex:"hql_tfactor|hql_backwy|zgemm_dmma" -c 3 -o /tmp/k45 python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/k45_ncu.log 2>&1
ncu -i /tmp/k45.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/k45_src.csv 2>/dev/null
ls -la gpurun_out/k45_src.csv
