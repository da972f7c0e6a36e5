"""In-graph cost of a dependent chain of small kernels (the encoder regime: M = 2560 rows)."""
import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
dev = "cuda"
M, D = 2560, 256
x = torch.randn(M, D, device=dev).bfloat16(); y = torch.empty_like(x)
g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
st = torch.empty(M, 2, device=dev)
w = torch.randn(D, D, device=dev).bfloat16(); w3 = torch.randn(768, D, device=dev).bfloat16()
o3 = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
tiny = torch.zeros(64, device=dev); tiny_b = torch.zeros(64, device=dev, dtype=torch.bfloat16)

def timed(fn, n=28, iters=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    for _ in range(3): gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / n * 1e3

print("cast 64 elements (launch floor)  us/launch", timed(lambda: ops.cast_bf16(tiny, tiny_b)))
print("layernorm fwd 2560x256           us/launch", timed(lambda: ops.layernorm_fwd(x, g, b, out=y, stats=st)))
print("gemm 2560x256x256                us/launch", timed(lambda: ops.gemm(x, w, out=y)))
print("gemm 2560x768x256                us/launch", timed(lambda: ops.gemm(x, w3, out=o3)))
qkv = torch.randn(M, 768, device=dev).bfloat16()
o, lse = ops.attention_fwd(qkv, 256, 10, 4, 64, 0.125)
print("attn small fwd                   us/launch", timed(lambda: ops.attention_fwd(qkv, 256, 10, 4, 64, 0.125, out=o, lse=lse)))
do = torch.randn(M, 256, device=dev).bfloat16(); dq = torch.empty_like(qkv)
print("attn small bwd                   us/launch", timed(lambda: ops.attention_bwd(qkv, o, do, lse, 256, 10, 4, 64, 0.125, dqkv=dq)))
