#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dinov2.py tests/test_raw_obs_gpu.py -x -q -m gpu > gpurun_out/r2d_new.log 2>&1; tail -25 gpurun_out/r2d_new.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench.err; tail -c 1500 gpurun_out/r2d_bench.log; tail -5 gpurun_out/r2d_bench.err
