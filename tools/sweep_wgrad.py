import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
def bench(fn, iters=10):
    for _ in range(2): fn()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
K = 49152
for (Mo, No) in [(256, 1024), (1024, 256), (256, 256), (768, 256)]:
    a = torch.randn(K, Mo, device="cuda").bfloat16(); b = torch.randn(K, No, device="cuda").bfloat16()
    out = torch.zeros(Mo, No, device="cuda")
    res = []
    for bn in (64, 128, 256):
        if No % bn: continue
        tiles = ((Mo + 127) // 128) * (No // bn)
        for target in (74, 148, 222, 296, 444, 592):
            s = max(1, target // tiles)
            if s > 768 // 2: continue
            us = bench(lambda: ops.gemm(a, b, mn_major=True, out=out, accumulate=True, splits=s, bn=bn))
            res.append((us, bn, s, tiles * s))
    res.sort()
    print(f"wgrad out {Mo}x{No} K={K}: best", [(f"{u:.1f}us", f"bn{bn}", f"s{s}", f"items{it}") for u, bn, s, it in res[:5]], "worst", f"{res[-1][0]:.1f}", flush=True)
