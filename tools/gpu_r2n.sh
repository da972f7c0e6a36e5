#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2n_all.log 2>&1; tail -5 gpurun_out/r2n_all.log
timeout 300 python bench.py --config 4 --steps 20 --warmup 5 > gpurun_out/r2n_cfg4.log 2>&1; tail -c 500 gpurun_out/r2n_cfg4.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.log 2> gpurun_out/r2n_bench.err; tail -c 300 gpurun_out/r2n_bench.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -2 gpurun_out/r2n_smoke.log
