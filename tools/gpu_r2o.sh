#!/bin/bash
mkdir -p gpurun_out
for k in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --profile 2>&1 | tail -1; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
