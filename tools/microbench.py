"""CUDA-event microbenchmarks of the decoder-shape kernels (B=256, n=192, D=256)."""
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
x = torch.randn(M, D, device=dev).bfloat16(); x2 = torch.randn(M, D, device=dev).bfloat16()
w_qkv = torch.randn(768, D, device=dev).bfloat16(); w_o = torch.randn(D, D, device=dev).bfloat16()
w1 = torch.randn(1024, D, device=dev).bfloat16(); w2 = torch.randn(D, 1024, device=dev).bfloat16()
b256 = torch.randn(D, device=dev); b1024 = torch.randn(1024, device=dev)
qkv = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
h = torch.randn(M, 1024, device=dev).bfloat16(); aux = torch.randn(M, 1024, device=dev).bfloat16()
out256 = torch.empty(M, D, device=dev, dtype=torch.bfloat16); out1024 = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(1024, device=dev)
print("qkv            ", t(lambda: ops.gemm(x, w_qkv, out=qkv)))
print("out-proj+res   ", t(lambda: ops.gemm(x, w_o, bias=b256, residual=x2, out=out256)))
print("ff1 plain      ", t(lambda: ops.gemm(x, w1, bias=b1024, out=out1024)))
print("ff1 gelu       ", t(lambda: ops.gemm(x, w1, bias=b1024, act=ops.GELU_FWD, out=out1024)))
print("ff1 gelu+aux   ", t(lambda: ops.gemm(x, w1, bias=b1024, act=ops.GELU_FWD, aux_out=aux, out=out1024)))
print("ff2+res        ", t(lambda: ops.gemm(h, w2, bias=b256, residual=x2, out=out256)))
print("dff2 plain     ", t(lambda: ops.gemm(x, w2.T.contiguous(), out=out1024)))
print("dff2 *aux      ", t(lambda: ops.gemm(x, w2.T.contiguous(), act=ops.GELU_BWD, aux_in=aux, out=out1024)))
print("dff2 *aux+cs   ", t(lambda: ops.gemm(x, w2.T.contiguous(), act=ops.GELU_BWD, aux_in=aux, out=out1024, colsum_out=cs)))
print("dff1 (K=1024)  ", t(lambda: ops.gemm(h, w1.T.contiguous(), out=out256)))
g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
dg, db, dc = (torch.zeros(D, device=dev) for _ in range(3))
y, st = ops.layernorm_fwd(x, g, b)
print("ln fwd         ", t(lambda: ops.layernorm_fwd(x, g, b, out=y, stats=st)))
print("ln bwd         ", t(lambda: ops.layernorm_bwd(x2, x, st, g, dgamma=dg, dbeta=db, skip=x2, dx_colsum=dc, dx=out256)))
qk = torch.randn(M, 768, device=dev).bfloat16()
o, lse = ops.attention_fwd(qk, 256, 192, 4, 64, 0.125)
print("attn fwd       ", t(lambda: ops.attention_fwd(qk, 256, 192, 4, 64, 0.125, out=o, lse=lse)))
dq = torch.empty_like(qk)
print("attn bwd       ", t(lambda: ops.attention_bwd(qk, o, x2, lse, 256, 192, 4, 64, 0.125, dqkv=dq)))
gw = torch.zeros(1024, 256, device=dev)
print("wgrad 1024x256 ", t(lambda: ops.gemm(h, x, mn_major=True, out=gw, accumulate=True, splits=18, bn=256)))
print("colsum 1024    ", t(lambda: ops.colsum(h, cs)))
qs = torch.randn(2560, 768, device=dev).bfloat16(); ds = torch.randn(2560, 256, device=dev).bfloat16()
os_, ls = ops.attention_fwd(qs, 256, 10, 4, 64, 0.125)
print("attn small fwd ", t(lambda: ops.attention_fwd(qs, 256, 10, 4, 64, 0.125, out=os_, lse=ls)))
print("attn small bwd ", t(lambda: ops.attention_bwd(qs, os_, ds, ls, 256, 10, 4, 64, 0.125)))
xs = torch.randn(2560, 256, device=dev).bfloat16(); outs = torch.empty(2560, 768, device=dev, dtype=torch.bfloat16)
print("enc qkv gemm   ", t(lambda: ops.gemm(xs, w_qkv, out=outs)))
ys, sts = ops.layernorm_fwd(xs, g, b)
print("enc ln fwd     ", t(lambda: ops.layernorm_fwd(xs, g, b, out=ys, stats=sts)))
