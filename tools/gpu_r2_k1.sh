python tools/eigh_check.py 64 96 83 72 49 50 65 60 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/hs_c5.json 2> gpurun_out/hs_c5.err; tail -c 400 gpurun_out/hs_c5.err
python -c "
import json;d=json.load(open('gpurun_out/hs_c5.json'));print(d['ms_per_step'],d['kernel_ms_per_step'],d.get('parity'))"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hql_tridiag" -c 4 --csv --log-file gpurun_out/hs_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/hs_ncu_launch.log 2>&1
grep -o 'hql_tridiag_[a-z]*_kernel<[0-9]*>.*' gpurun_out/hs_launches.csv | sed 's/(int.*gpu__time_duration.sum//' 
