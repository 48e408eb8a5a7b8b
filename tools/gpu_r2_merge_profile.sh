ncu --set full --clock-control none --import-source on -k regex:"tdc_merge" -c 1 -o /tmp/mg python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/mg_ncu.log 2>&1
ncu -i /tmp/mg.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/mg_src.csv 2>/dev/null
