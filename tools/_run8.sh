mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/r1e_bench_2gpu.json 2> gpurun_out/r1e_bench_2gpu.err; tail -c 1500 gpurun_out/r1e_bench_2gpu.json; tail -3 gpurun_out/r1e_bench_2gpu.err
timeout 150 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
