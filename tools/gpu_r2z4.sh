#!/bin/bash
# four-GPU line of the final code
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 > gpurun_out/r2z_bench_4gpu.json 2> gpurun_out/r2z_bench_4gpu.err; tail -c 200 gpurun_out/r2z_bench_4gpu.json; tail -2 gpurun_out/r2z_bench_4gpu.err
