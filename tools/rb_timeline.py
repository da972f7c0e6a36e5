"""Event timeline of the fused feed-forward block kernel (CTA 0), profile build only:
    python tools/build_variant.py rbprof -DM3L_RB_PROFILE
    M3L_B200_LIB=$PWD/m3l_b200/lib/variant_rbprof.so python tools/rb_timeline.py [save]"""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops, _lib
lib = _lib.load()
save = len(sys.argv) > 1 and sys.argv[1] == "save"
M, hidden, D = 49152, 1024, 256
dev = "cuda"
x = torch.randn(M, D, device=dev).bfloat16()
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
w1 = (torch.randn(hidden, D, device=dev) * 0.05).bfloat16(); b1 = torch.zeros(hidden, device=dev)
w2 = (torch.randn(D, hidden, device=dev) * 0.05).bfloat16(); b2 = torch.zeros(D, device=dev)
buf = (C.c_longlong * 2048)()
for _ in range(3):
    ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2, save=save, out=(x.clone() if save else x), out_has_x=save)
    lib.m3l_debug_rb_prof(buf, 2048)
nc = hidden // 128
ev = lambda role, i: buf[role * 512 + i]
t0 = min(v for v in buf if v > 0)
rows = []
def item_name(it):
    if it == 2 * nc - 1: return f"G2({nc-1})"
    if it < 2: return f"G1({it})"
    return f"G2({it//2-1})" if it % 2 == 0 else f"G1({(it+1)//2})"
for tile in range(3):
    for it in range(2 * nc):
        k = tile * 2 * nc + it
        if ev(0, k): rows.append((ev(0, k) - t0, f"TMA  t{tile} weights issued for {item_name(it)}"))
        if ev(1, 2 * k): rows.append((ev(1, 2 * k) - t0, f"MMA  t{tile} {item_name(it)} issue start"))
        if ev(1, 2 * k + 1): rows.append((ev(1, 2 * k + 1) - t0, f"MMA  t{tile} {item_name(it)} issued"))
    names = ["LN start (tile landed)", "LN done", "out start (acc2 full)", "out done"]
    for e in range(4):
        if ev(2, tile * 4 + e): rows.append((ev(2, tile * 4 + e) - t0, f"ROW  t{tile} {names[e]}"))
    for c in range(nc):
        k = tile * nc + c
        gn = ["acc1 full seen", "tmem ld done", "math done", "h stored + arrived"]
        for e in range(4):
            if ev(3, k * 4 + e): rows.append((ev(3, k * 4 + e) - t0, f"GELU t{tile} c{c} {gn[e]}"))
rows.sort()
for t, nm in rows:
    print(f"{t:9d}  {nm}")
