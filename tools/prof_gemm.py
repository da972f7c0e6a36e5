"""One launch each of the decoder GEMM shapes after a warm-up round (for ncu --set full captures).
Order of the profiled launches: qkv, out-proj+res, ff1 gelu+aux, ff2+res, dff2*aux+colsum, dff1, wgrad."""
import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops
M, D = 49152, 256
dev = "cuda"
x = torch.randn(M, D, device=dev).bfloat16(); x2 = torch.randn(M, D, device=dev).bfloat16()
w_qkv = torch.randn(768, D, device=dev).bfloat16(); w_o = torch.randn(D, D, device=dev).bfloat16()
w1 = torch.randn(1024, D, device=dev).bfloat16(); w2 = torch.randn(D, 1024, device=dev).bfloat16()
w2t = w2.T.contiguous(); w1t = w1.T.contiguous()
b256 = torch.randn(D, device=dev); b1024 = torch.randn(1024, device=dev)
qkv = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
h = torch.randn(M, 1024, device=dev).bfloat16(); aux = torch.randn(M, 1024, device=dev).bfloat16()
out256 = torch.empty(M, D, device=dev, dtype=torch.bfloat16); out1024 = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(1024, device=dev)
gw = torch.zeros(1024, 256, device=dev)
for _ in range(2):
    ops.gemm(x, w_qkv, out=qkv)
    ops.gemm(x, w_o, bias=b256, residual=x2, out=out256)
    ops.gemm(x, w1, bias=b1024, act=ops.GELU_FWD, aux_out=aux, out=out1024)
    ops.gemm(h, w2, bias=b256, residual=x2, out=out256)
    ops.gemm(x, w2t, act=ops.GELU_BWD, aux_in=aux, out=out1024, colsum_out=cs)
    ops.gemm(h, w1t, out=out256)
    ops.gemm(h, x, mn_major=True, out=gw, accumulate=True, splits=18, bn=256)
torch.cuda.synchronize()
print("done")
