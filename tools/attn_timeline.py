"""Event timeline of the attention backward kernel (CTA 0), profile build only:
    python tools/build_variant.py prof -DM3L_ATTN_PROFILE
    M3L_B200_LIB=$PWD/m3l_b200/lib/variant_prof.so M3L_ATTN_PROF=1 python tools/attn_timeline.py"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops, _lib
lib = _lib.load()
B, n, H = 256, 192, 4
M = B * n
dev = "cuda"
qkv = torch.randn(M, 768, device=dev).bfloat16(); do = torch.randn(M, 256, device=dev).bfloat16()
o, lse = ops.attention_fwd(qkv, B, n, H, 64, 0.125)
delta = (do.float() * o.float()).reshape(M, H, 64).sum(-1).contiguous()
dq = torch.empty_like(qkv)
buf = (C.c_longlong * 2048)()
for _ in range(2):
    ops.attention_bwd(qkv, o, do, lse, B, n, H, 64, 0.125, dqkv=dq, delta=delta); lib.m3l_debug_attn_prof(buf, 2048)
ev = lambda role, g, e: buf[64 + role * 512 + g * 8 + e]
t0 = min(v for v in (ev(r, 0, e) for r in range(3) for e in range(8)) if v > 0)
names = {0: ["sdp_free seen", "S,dP(next) issued", "pds_full seen", "dV,dK,dQ issued"],
         1: ["item start", "sdp_full seen", "LDTM+math done", "slabs free", "STS done", "dKV drained", "item dKV", "item dQ"]}
for g in range(12):
    rows = []
    for r in (0, 1, 2):
        for e in range(8):
            v = ev(r, g, e)
            if v > 0:
                nm = names[0][e] if r == 0 else names[1][e]
                rows.append((v - t0, f"{'MMA ' if r == 0 else 'wg' + str(r - 1) + ' '}{nm}"))
    rows.sort()
    print(f"--- step {g}")
    for t, nm in rows:
        print(f"   {t:8d}  {nm}")
