#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_all.log 2>&1; tail -5 gpurun_out/r2s_all.log
for k in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --profile 2>&1 | tail -1; done
M3L_FUSED_LN_BWD=0 timeout 300 python bench.py --steps 20 --warmup 5 --profile 2>&1 | tail -1
timeout 300 python tools/step_breakdown.py > gpurun_out/r2s_breakdown.log 2>&1; cat gpurun_out/r2s_breakdown.log
