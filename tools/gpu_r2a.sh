#!/bin/bash
# round-2 session-2 baseline: tests, microbench, bench, breakdown, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1; tail -4 gpurun_out/r2a_tests.log
timeout 300 python tools/bench_mlp.py > gpurun_out/r2a_mlp.log 2>&1; tail -5 gpurun_out/r2a_mlp.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.log
timeout 300 python tools/step_breakdown.py > gpurun_out/r2a_breakdown.log 2>&1; cat gpurun_out/r2a_breakdown.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --profile > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log
