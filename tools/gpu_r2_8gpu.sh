set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_8gpu.json 2> gpurun_out/r2_b_8gpu.err; tail -5 gpurun_out/r2_b_8gpu.err
python tools/show_bench.py gpurun_out/r2_bench_c5_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c5_4gpu.json 2> gpurun_out/r2_b_4gpu.err
python tools/show_bench.py gpurun_out/r2_bench_c5_4gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c5_2gpu.json 2> gpurun_out/r2_b_2gpu.err
python tools/show_bench.py gpurun_out/r2_bench_c5_2gpu.json
