#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "ln_mlp or epilogues" > gpurun_out/g_ops.log 2>&1
echo "ops rc=$?" >> gpurun_out/g_ops.log
tail -4 gpurun_out/g_ops.log
timeout 300 python tools/bench_mlp.py > gpurun_out/g_mlp.log 2>&1; tail -5 gpurun_out/g_mlp.log
python tools/build_variant.py rbprof -DM3L_RB_PROFILE > gpurun_out/g_build.log 2>&1
M3L_B200_LIB=$PWD/m3l_b200/lib/variant_rbprof.so timeout 300 python tools/rb_timeline.py > gpurun_out/g_timeline.log 2>&1
M3L_B200_LIB=$PWD/m3l_b200/lib/variant_rbprof.so timeout 300 python tools/rb_timeline.py save > gpurun_out/g_timeline_save.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/g_all.log 2>&1; tail -3 gpurun_out/g_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.log 2> gpurun_out/g_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/g_bench.log').read().strip().splitlines()[-1]);print('bench ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])"
