#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_torch_ops.py tests/test_raw_obs_gpu.py -x -q -m gpu > gpurun_out/r2c_new.log 2>&1; tail -25 gpurun_out/r2c_new.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2c_all.log 2>&1; tail -8 gpurun_out/r2c_all.log
timeout 300 python bench.py --steps 20 --warmup 5 --profile > gpurun_out/r2c_bench.log 2>&1; tail -1 gpurun_out/r2c_bench.log
