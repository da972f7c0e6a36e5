#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "layernorm_backward_epilogue or gemm" > gpurun_out/r2r_tests.log 2>&1; tail -15 gpurun_out/r2r_tests.log
