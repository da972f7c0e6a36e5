#!/bin/bash
mkdir -p gpurun_out
python tools/build_variant.py rbprof -DM3L_RB_PROFILE > gpurun_out/r2b_build.log 2>&1 || { tail -5 gpurun_out/r2b_build.log; exit 1; }
M3L_B200_LIB=$PWD/m3l_b200/lib/variant_rbprof.so timeout 300 python tools/rb_timeline.py > gpurun_out/r2b_timeline.log 2>&1
M3L_B200_LIB=$PWD/m3l_b200/lib/variant_rbprof.so timeout 300 python tools/rb_timeline.py save > gpurun_out/r2b_timeline_save.log 2>&1
cat > /tmp/prof_mlp.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
dev = "cuda"; M, hidden, D = 49152, 1024, 256
x = torch.randn(M, D, device=dev).bfloat16()
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
w1 = (torch.randn(hidden, D, device=dev) * 0.05).bfloat16(); b1 = torch.zeros(hidden, device=dev)
w2 = (torch.randn(D, hidden, device=dev) * 0.05).bfloat16(); b2 = torch.zeros(D, device=dev)
buf = x.clone()
for _ in range(2):
    ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2, out=x)
    ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2, save=True, out=buf, out_has_x=True)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_mlp -c 4 -o gpurun_out/prof_rb_r2b -f python /tmp/prof_mlp.py > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
