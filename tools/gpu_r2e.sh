#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vtmae_gpu.py -x -q -m gpu -k "joint" > gpurun_out/r2e_new.log 2>&1; tail -30 gpurun_out/r2e_new.log
