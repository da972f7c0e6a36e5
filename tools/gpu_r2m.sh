#!/bin/bash
# N-GPU pass (N = $1): headline + secondary configs, tight timeouts
N=$1
mkdir -p gpurun_out
for cfg in $2; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$cfg bench.py --gpus $N --config $cfg --steps 20 --warmup 5 > gpurun_out/r2m_cfg${cfg}_n${N}.log 2> gpurun_out/r2m_cfg${cfg}_n${N}.err
  echo "cfg $cfg N=$N rc=$?"; tail -c 400 gpurun_out/r2m_cfg${cfg}_n${N}.log; echo
done
