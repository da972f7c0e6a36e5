#!/bin/bash
# ncu --set full of the per-kernel-table entries added at the end of round 2 (the other entries: r02_ncu_top_kernels.txt)
mkdir -p gpurun_out
M3L_PER_KERNEL_ONLY=dgrad_ff1_ln,dgrad_qkv_ln,out_proj_dgrad timeout 600 ncu --set full --clock-control none -k regex:'gemm_bf16' -c 9 -o /tmp/r2x_new -f python tools/ncu_top_kernels.py > gpurun_out/r2x_ncu_new.log 2>&1; tail -2 gpurun_out/r2x_ncu_new.log
ncu -i /tmp/r2x_new.ncu-rep --page raw --csv > gpurun_out/r2x_new_raw.csv 2>/dev/null; wc -c gpurun_out/r2x_new_raw.csv
