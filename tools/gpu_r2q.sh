#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.log 2> gpurun_out/r2q_bench.err; tail -c 200 gpurun_out/r2q_bench.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2q_ref.log 2>&1; tail -c 300 gpurun_out/r2q_ref.log
timeout 300 python tools/bench_configs.py cfg5 > gpurun_out/r2q_cfg5.log 2>&1; cat gpurun_out/r2q_cfg5.log | cut -c1-300
timeout 300 python bench.py --config 3 --steps 20 --warmup 5 > gpurun_out/r2q_cfg3.log 2>&1; tail -c 200 gpurun_out/r2q_cfg3.log
timeout 300 python bench.py --config 4 --steps 20 --warmup 5 > gpurun_out/r2q_cfg4.log 2>&1; tail -c 200 gpurun_out/r2q_cfg4.log
timeout 300 python tools/eager_latency.py > gpurun_out/r2q_latency.log 2>&1; tail -9 gpurun_out/r2q_latency.log
