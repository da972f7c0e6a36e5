#!/bin/bash
# 2-GPU verification with tight timeouts: dp worker, one-graph bench exit behaviour
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_worker.py > gpurun_out/r2i_dpworker.log 2>&1; echo "dp_worker rc=$?"; grep -a "dp_worker\|DP_WORKER\|Error\|error" gpurun_out/r2i_dpworker.log | tail -8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.log 2> gpurun_out/r2i_bench_n2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2i_bench_n2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_ref_n2.log 2>&1; echo "ref rc=$?"; tail -c 300 gpurun_out/r2i_ref_n2.log
timeout 300 python -m pytest tests/test_vtt_dino_gpu.py tests/test_vtmae_gpu.py -x -q -m gpu -k "vtdino or ppo_mae" > gpurun_out/r2i_new.log 2>&1; tail -5 gpurun_out/r2i_new.log
