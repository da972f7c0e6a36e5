import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
B, n, H = 256, 192, 4
M = B * n
qkv = torch.randn(M, 768, device=dev).bfloat16()
do = torch.randn(M, 256, device=dev).bfloat16()
o, lse = ops.attention_fwd(qkv, B, n, H, 64, 0.125)
buf = (C.c_longlong * 16)()
for _ in range(2):
    ops.attention_fwd(qkv, B, n, H, 64, 0.125, out=o, lse=lse); lib.m3l_debug_attn_prof(buf, 16)
t = max(buf[5], 1)
print(f"fwd  tiles(warp2)={buf[5]} per tile [clk]: wait S {buf[0]/t:.0f} | max pass {buf[1]/t:.0f} | exp pass {buf[2]/t:.0f} | wait O {buf[3]/t:.0f} | epilogue {buf[4]/t:.0f} | total/tile {buf[6]/t:.0f}")
dq = torch.empty_like(qkv)
for _ in range(2):
    ops.attention_bwd(qkv, o, do, lse, B, n, H, 64, 0.125, dqkv=dq); lib.m3l_debug_attn_prof(buf, 16)
t = max(buf[6], 1)
print(f"bwd  steps={buf[6]} per step [clk]: delta {buf[0]/t:.0f} | wait S,dP {buf[1]/t:.0f} | P/dS work {buf[2]/t:.0f} | wait dKV {buf[3]/t:.0f} | dKV epi {buf[4]/t:.0f} | dQ epi {buf[5]/t:.0f} | wait slabs free {buf[8]/t:.0f} | STS {buf[9]/t:.0f} | total/step {buf[7]/t:.0f}")
print(f"bwd MMA thread per step [clk]: wait loads {buf[10]/t:.0f} | wait sdp_free {buf[11]/t:.0f} | issue SdP {buf[12]/t:.0f} | wait pds_full {buf[13]/t:.0f} | wait acc free {buf[14]/t:.0f} | issue dV,dK,dQ {buf[15]/t:.0f}")
def tm(fn, it=20):
    for _ in range(3): fn()
    s0 = torch.cuda.Event(enable_timing=True); e0 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s0.record()
    for _ in range(it): fn()
    e0.record(); torch.cuda.synchronize(); return s0.elapsed_time(e0) / it * 1e3
print("attn fwd us", tm(lambda: ops.attention_fwd(qkv, B, n, H, 64, 0.125, out=o, lse=lse)))
print("attn bwd us", tm(lambda: ops.attention_bwd(qkv, o, do, lse, B, n, H, 64, 0.125, dqkv=dq)))
