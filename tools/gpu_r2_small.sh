python tools/eigh_check.py 9 12 16 17 24 25 31 32 96 2>&1 | tail -10
for w in c2 c3; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/hsw_$w.json 2> gpurun_out/hsw_$w.err; tail -c 300 gpurun_out/hsw_$w.err; python -c "
import json;d=json.load(open('gpurun_out/hsw_$w.json'));print('$w', d['ms_per_step'],d['kernel_ms_per_step'])"; done
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/hs_c5.json 2> gpurun_out/hs_c5.err; python -c "
import json;d=json.load(open('gpurun_out/hs_c5.json'));print('c5', d['ms_per_step'],d['kernel_ms_per_step'])"
