#!/bin/bash
# two-GPU check of the final code: bit-identical replicas (tests/dp_worker.py) and the headline bench line at N = 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bench_shapes_gpu.py -q -m gpu -k two_ranks > gpurun_out/r2z_dp_test.log 2>&1; tail -2 gpurun_out/r2z_dp_test.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2z_bench_2gpu.json 2> gpurun_out/r2z_bench_2gpu.err; tail -c 300 gpurun_out/r2z_bench_2gpu.json; tail -2 gpurun_out/r2z_bench_2gpu.err
