import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
def t(fn, it=20):
    for _ in range(3): fn()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it * 1e3
M, D = 49152, 256
dev = "cuda"
x = torch.randn(M, D, device=dev).bfloat16()
w_qkv = torch.randn(768, D, device=dev).bfloat16(); w_o = torch.randn(D, D, device=dev).bfloat16()
w1 = torch.randn(1024, D, device=dev).bfloat16()
qkv = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
out256 = torch.empty(M, D, device=dev, dtype=torch.bfloat16); out1024 = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
print("qkv   ", t(lambda: ops.gemm(x, w_qkv, out=qkv)))
print("out   ", t(lambda: ops.gemm(x, w_o, out=out256)))
print("ff1   ", t(lambda: ops.gemm(x, w1, out=out1024)))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10): ops.gemm(x, w_qkv, out=qkv)
print("qkv graph x10 per-launch", t(lambda: g.replay(), it=5) / 10)
