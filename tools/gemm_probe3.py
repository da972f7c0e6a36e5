import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
def prof(label, fn):
    buf = (C.c_longlong * 8)()
    fn(); lib.m3l_debug_gemm_prof(buf, 8)
    fn(); lib.m3l_debug_gemm_prof(buf, 8)
    t = max(buf[6], 1)
    print(f"{label:22s} tiles(CTA0)={buf[6]:3d} per tile [clk]: producer wait empty {buf[0]/t:7.0f} | mma: wait tmem_empty {buf[1]/t:7.0f} wait full {buf[2]/t:7.0f} whole {buf[3]/t:7.0f} | epi: wait tmem_full {buf[4]/t:7.0f} work {buf[5]/t:7.0f}")
M, D = 49152, 256
x = torch.randn(M, D, device=dev).bfloat16(); x2 = torch.randn(M, D, device=dev).bfloat16()
w_qkv = torch.randn(768, D, device=dev).bfloat16()
w1 = torch.randn(1024, D, device=dev).bfloat16(); w2 = torch.randn(D, 1024, device=dev).bfloat16()
b1024 = torch.randn(1024, device=dev); b256 = torch.randn(D, device=dev)
h = torch.randn(M, 1024, device=dev).bfloat16(); aux = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16)
qkv = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
out1024 = torch.empty(M, 1024, device=dev, dtype=torch.bfloat16); out256 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
prof("qkv", lambda: ops.gemm(x, w_qkv, out=qkv))
prof("ff1 gelu+aux", lambda: ops.gemm(x, w1, bias=b1024, act=ops.GELU_FWD, aux_out=aux, out=out1024))
prof("ff2+res (K=1024)", lambda: ops.gemm(h, w2, bias=b256, residual=x2, out=out256))
