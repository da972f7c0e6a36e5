#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_all.log 2>&1; tail -6 gpurun_out/r2f_all.log
timeout 600 python tools/eager_latency.py > gpurun_out/r2f_latency.log 2>&1; tail -12 gpurun_out/r2f_latency.log
timeout 300 python bench.py --config 3 --steps 20 --warmup 5 > gpurun_out/r2f_cfg3.log 2>&1; tail -c 900 gpurun_out/r2f_cfg3.log
timeout 300 python bench.py --config 4 --steps 20 --warmup 5 > gpurun_out/r2f_cfg4.log 2>&1; tail -c 900 gpurun_out/r2f_cfg4.log
