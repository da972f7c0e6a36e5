#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_bench_shapes_gpu.py -x -q -m gpu -k "ln_mlp or full_step or train_step or bench" > gpurun_out/r2l_tests.log 2>&1; tail -5 gpurun_out/r2l_tests.log
timeout 300 python tools/bench_mlp.py > gpurun_out/r2l_mlp.log 2>&1; tail -4 gpurun_out/r2l_mlp.log
timeout 300 python bench.py --steps 20 --warmup 5 --profile > gpurun_out/r2l_bench.log 2>&1; tail -1 gpurun_out/r2l_bench.log
