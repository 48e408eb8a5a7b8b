cp muspinsim_b200/csrc/libmusim.so /tmp/base.so
for m in 4 5 6; do
if [ $m != 4 ]; then cp tools/scratch/libmusim_hsw$m.so muspinsim_b200/csrc/libmusim.so; fi
python bench.py --workload c3 --steps 2 --warmup 2 --no-cpu --option tridiag_hsw=1 > gpurun_out/hsw_c3_$m.json 2> gpurun_out/hsw_c3_$m.err || tail -3 gpurun_out/hsw_c3_$m.err
python -c "
import json;d=json.load(open('gpurun_out/hsw_c3_$m.json'));print('minb $m', d['ms_per_step'],d['kernel_ms_per_step']['eigh_tridiag'])"
done
cp /tmp/base.so muspinsim_b200/csrc/libmusim.so
