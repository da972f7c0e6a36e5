"""Decoder-shape launches of the HBM-bound kernels (for ncu captures)."""
import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
M, D = 49152, 256
x = torch.randn(M, D, device="cuda").bfloat16()
dy = torch.randn(M, D, device="cuda").bfloat16()
skip = torch.randn(M, D, device="cuda").bfloat16()
g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
dg, db, dc = (torch.zeros(D, device="cuda") for _ in range(3))
big = torch.randn(M, 1024, device="cuda").bfloat16()
out1024 = torch.zeros(1024, device="cuda")
img = torch.rand(256, 12, 64, 64, device="cuda")
ps = ops.make_patch_source([img], 8, 8, 0)
idx = torch.stack([torch.randperm(64)[:60] for _ in range(256)]).cuda()
pred = torch.randn(256 * 60, 768, device="cuda")
acc = torch.zeros(1, device="cuda")
qkv = torch.randn(M, 768, device="cuda").bfloat16()
for _ in range(3):
    y, st = ops.layernorm_fwd(x, g, b)
    ops.layernorm_bwd(dy, x, st, g, dgamma=dg, dbeta=db, skip=skip, dx_colsum=dc)
    ops.colsum(big, out1024)
    ops.mse_loss(ps, 256, 60, pred, 1e-6, acc, tok_idx=idx)
    o, lse = ops.attention_fwd(qkv, 256, 192, 4, 64, 0.125)
    ops.attention_bwd(qkv, o, dy, lse, 256, 192, 4, 64, 0.125)
torch.cuda.synchronize()
print("done")
