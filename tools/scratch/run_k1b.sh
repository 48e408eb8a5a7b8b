python tools/scratch/hs_dbg.py 64 96 83 50 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/hs_c5.json 2> gpurun_out/hs_c5.err; tail -c 300 gpurun_out/hs_c5.err
python -c "
import json;d=json.load(open('gpurun_out/hs_c5.json'));print('base', d['ms_per_step'],d['kernel_ms_per_step']['eigh_tridiag'])"
cp muspinsim_b200/csrc/libmusim.so /tmp/libmusim_base.so
cp tools/scratch/libmusim_dmmarows.so muspinsim_b200/csrc/libmusim.so
python tools/scratch/hs_dbg.py 64 96 83 50 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/hs_c5b.json 2> gpurun_out/hs_c5b.err; tail -c 300 gpurun_out/hs_c5b.err
python -c "
import json;d=json.load(open('gpurun_out/hs_c5b.json'));print('dmma rows', d['ms_per_step'],d['kernel_ms_per_step']['eigh_tridiag'])"
