#!/bin/bash
# 2-GPU pass: multi-rank replica test, headline bench and the secondary configs at N=2
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_vtt_dino_gpu.py tests/test_vtmae_gpu.py tests/test_bench_shapes_gpu.py -x -q -m gpu -k "vtdino or ppo_mae or two_ranks" > gpurun_out/r2g_new.log 2>&1; tail -15 gpurun_out/r2g_new.log
for cfg in 2 3 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --config $cfg --steps 20 --warmup 5 > gpurun_out/r2g_bench_cfg${cfg}_n2.log 2> gpurun_out/r2g_bench_cfg${cfg}_n2.err
  echo "cfg $cfg rc=$?"; tail -c 700 gpurun_out/r2g_bench_cfg${cfg}_n2.log; tail -3 gpurun_out/r2g_bench_cfg${cfg}_n2.err
done
M3L_DP_ONE_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 5 --profile > gpurun_out/r2g_bench_n2_multigraph.log 2>&1; tail -1 gpurun_out/r2g_bench_n2_multigraph.log
