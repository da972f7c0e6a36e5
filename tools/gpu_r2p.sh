#!/bin/bash
# final evidence pass of the round (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2p_all.log 2>&1; tail -4 gpurun_out/r2p_all.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|ln_mlp|attn_|ln_bwd_pipe|ln_fwd_pipe' -c 60 -o gpurun_out/r2p_top -f python tools/ncu_top_kernels.py > gpurun_out/r2p_ncu_top.log 2>&1; tail -2 gpurun_out/r2p_ncu_top.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --profile > gpurun_out/r2p_ncu_list.log 2>&1
timeout 300 python tools/step_breakdown.py > gpurun_out/r2p_breakdown.log 2>&1; cat gpurun_out/r2p_breakdown.log
timeout 300 python tools/attn_probe.py 2>&1 | tail -3 > gpurun_out/r2p_attn_times.log; cat gpurun_out/r2p_attn_times.log
