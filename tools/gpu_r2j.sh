#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" > gpurun_out/r2j_attn.log 2>&1; tail -4 gpurun_out/r2j_attn.log
timeout 300 python tools/attn_probe.py 2>&1 | tail -3 > gpurun_out/r2j_probe.log; cat gpurun_out/r2j_probe.log
timeout 300 python bench.py --steps 20 --warmup 5 --profile > gpurun_out/r2j_bench.log 2>&1; tail -1 gpurun_out/r2j_bench.log
