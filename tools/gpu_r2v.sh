#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; tail -3 gpurun_out/r2v_tests.log
timeout 100 python tools/mse_probe.py
M3L_MSE_FAST=0 timeout 100 python tools/mse_probe.py
timeout 300 python bench.py --steps 20 --warmup 5 --profile 2>&1 | tail -1
M3L_MSE_FAST=0 timeout 300 python bench.py --steps 20 --warmup 5 --profile 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --profile 2>&1 | tail -1
timeout 300 python tools/step_breakdown.py 2>&1 | head -3
