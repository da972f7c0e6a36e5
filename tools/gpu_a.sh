#!/bin/bash
# round-2 GPU pass A: fused feed-forward block — parity first, then timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "ln_mlp or weight_stationary" > gpurun_out/a_ops.log 2>&1
echo "ops rc=$?" >> gpurun_out/a_ops.log
tail -15 gpurun_out/a_ops.log
timeout 300 python tools/bench_mlp.py > gpurun_out/a_mlp.log 2>&1; tail -8 gpurun_out/a_mlp.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_all.log 2>&1; tail -8 gpurun_out/a_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.log 2> gpurun_out/a_bench.err; tail -2 gpurun_out/a_bench.log
M3L_FUSED_MLP=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_unfused.log 2> gpurun_out/a_bench_unfused.err; tail -2 gpurun_out/a_bench_unfused.log
