import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
def t(fn, it=10):
    for _ in range(3): fn()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it * 1e3
dev = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
w = torch.randn(N, K, device=dev).bfloat16()
for tiles in (1, 2, 4, 8, 16, 32):
    M = 148 * 128 * tiles
    x = torch.randn(M, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, w, out=out)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): ops.gemm(x, w, out=out)
    print(f"N={N} K={K} m-tiles/CTA={tiles:3d}  per-launch {t(lambda: g.replay()) / 10:8.2f} us")
