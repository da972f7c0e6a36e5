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
delta = (do.float() * o.float()).reshape(M, H, 64).sum(-1).contiguous()
buf = (C.c_longlong * 32)()
for _ in range(2):
    ops.attention_bwd(qkv, o, do, lse, B, n, H, 64, 0.125, dqkv=dq, delta=delta); lib.m3l_debug_attn_prof(buf, 32)
names = ["stats", "wait S,dP", "LDTM+math", "wait slabs", "STS+arrive", "wait dKV", "dKV epi", "dQ epi"]
for o_, w_ in ((0, "warp4 (group 0, rows 0-31)"), (16, "warp8 (group 1, rows 0-31)")):
    t_ = max(buf[o_ + 8], 1)
    print(f"bwd {w_} per step [clk]: " + " | ".join(f"{nm} {buf[o_ + k] / t_:.0f}" for k, nm in enumerate(names)) + f" | total {buf[o_ + 9] / t_:.0f}")
t_ = max(buf[8], 1)
print(f"bwd MMA thread per step [clk]: wait sdp_free {buf[26]/t_:.0f} | S,dP issue->complete {buf[27]/t_:.0f} | wait pds_full {buf[28]/t_:.0f} | dV,dK,dQ drain {buf[29]/t_:.0f}")
def tm(fn, it=20):
    for _ in range(3): fn()
    s0 = torch.cuda.Event(enable_timing=True); e0 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s0.record()
    for _ in range(it): fn()
    e0.record(); torch.cuda.synchronize(); return s0.elapsed_time(e0) / it * 1e3
print("attn fwd us", tm(lambda: ops.attention_fwd(qkv, B, n, H, 64, 0.125, out=o, lse=lse)))
print("attn bwd us (in-kernel delta)", tm(lambda: ops.attention_bwd(qkv, o, do, lse, B, n, H, 64, 0.125, dqkv=dq)))
print("attn bwd us (delta given)", tm(lambda: ops.attention_bwd(qkv, o, do, lse, B, n, H, 64, 0.125, dqkv=dq, delta=delta)))
