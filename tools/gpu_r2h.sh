#!/bin/bash
mkdir -p gpurun_out
python tools/build_variant.py prof -DM3L_ATTN_PROFILE > gpurun_out/r2h_build.log 2>&1 || { tail -5 gpurun_out/r2h_build.log; exit 1; }
M3L_B200_LIB=$PWD/m3l_b200/lib/variant_prof.so M3L_ATTN_PROF=1 timeout 300 python tools/attn_probe.py > gpurun_out/r2h_probe.log 2>&1; cat gpurun_out/r2h_probe.log
M3L_B200_LIB=$PWD/m3l_b200/lib/variant_prof.so M3L_ATTN_PROF=1 timeout 300 python tools/attn_timeline.py > gpurun_out/r2h_timeline.log 2>&1; head -150 gpurun_out/r2h_timeline.log
