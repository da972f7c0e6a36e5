#!/bin/bash
# final evidence pass (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2w_all.log 2>&1; tail -3 gpurun_out/r2w_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -c 300 gpurun_out/r2w_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2w_ref.json 2> gpurun_out/r2w_ref.err; tail -c 400 gpurun_out/r2w_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --profile > gpurun_out/r2w_ncu_list.log 2>&1
timeout 300 python tools/step_breakdown.py > gpurun_out/r2w_breakdown.log 2>&1; cat gpurun_out/r2w_breakdown.log
timeout 100 python tools/mse_probe.py > gpurun_out/r2w_mse.log 2>&1; cat gpurun_out/r2w_mse.log
