# Final round-2 bench lines of every workload (one B200), with the CPU arm's parity block.
set -x
for w in c5 c2 c3 c4; do python bench.py --workload $w > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_b_$w.err; tail -2 gpurun_out/r2_b_$w.err; done
python bench.py --general > gpurun_out/r2_bench_c5_general.json 2> gpurun_out/r2_b_c5g.err
python bench.py --workload c2 --general > gpurun_out/r2_bench_c2_general.json 2> gpurun_out/r2_b_c2g.err
python bench.py --n-orient 2500 --no-cpu > gpurun_out/r2_bench_c5_2500.json 2> gpurun_out/r2_b_c5_2500.err
for f in gpurun_out/r2_bench_c*.json; do python tools/show_bench.py $f; done
