"""Runs a few launches of the dominant GEMM shapes (for ncu --set full captures)."""
import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
M = 49152
a = torch.randn(M, 256, device="cuda").bfloat16()
w_qkv = torch.randn(768, 256, device="cuda").bfloat16()
w1 = torch.randn(1024, 256, device="cuda").bfloat16()
b1 = torch.randn(1024, device="cuda")
out = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
h = torch.empty(M, 1024, device="cuda", dtype=torch.bfloat16)
pre = torch.empty_like(h)
for _ in range(3):
    ops.gemm(a, w_qkv, out=out)
    ops.gemm(a, w1, out=h, bias=b1, act=ops.GELU_FWD, aux_out=pre)
torch.cuda.synchronize()
print("done")
