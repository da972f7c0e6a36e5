#!/bin/bash
# evidence pass: launch list, step breakdown, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 300 python tools/step_breakdown.py > gpurun_out/r2k_breakdown.log 2>&1; cat gpurun_out/r2k_breakdown.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --profile > gpurun_out/r2k_ncu_list.log 2>&1; tail -1 gpurun_out/r2k_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|ln_mlp|attn_|ln_bwd_pipe|ln_fwd_pipe' --launch-skip 0 -c 60 -o gpurun_out/r2k_top -f python tools/ncu_top_kernels.py > gpurun_out/r2k_ncu_top.log 2>&1; tail -3 gpurun_out/r2k_ncu_top.log
ls -la gpurun_out/r2k_top.ncu-rep
