"""Times the fused feed-forward block (csrc/rowblock.cu) against the unfused LayerNorm / FF1+GELU / FF2 kernels
at the decoder (M = 49152, hidden 1024) and encoder (M = 2560, hidden 512) shapes.  usage: python tools/bench_mlp.py"""
import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops

dev = "cuda"


def timeit(fn, warm=5, iters=30):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for M, hidden in [(49152, 1024), (2560, 512), (8192, 1024), (4096 * 192, 512)]:
    torch.manual_seed(0)
    D = 256
    x = torch.randn(M, D, device=dev).bfloat16()
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    w1 = (torch.randn(hidden, D, device=dev) * 0.05).bfloat16(); b1 = torch.zeros(hidden, device=dev)
    w2 = (torch.randn(D, hidden, device=dev) * 0.05).bfloat16(); b2 = torch.zeros(D, device=dev)

    def unfused(train):
        xn, st = ops.layernorm_fwd(x, gamma, beta, want_stats=train)
        pre = torch.empty((M, hidden), dtype=torch.bfloat16, device=dev) if train else None
        h = ops.gemm(xn, w1, bias=b1, act=ops.GELU_FWD, aux_out=pre)
        return ops.gemm(h, w2, bias=b2, residual=x)

    t_f0 = timeit(lambda: ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2))
    xi = x.clone()
    t_f = timeit(lambda: ops.ln_mlp_fwd(xi, gamma, beta, w1, b1, w2, b2, out=xi))
    buf = x.clone()
    t_fs = timeit(lambda: ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2, save=True, out=buf, out_has_x=True))
    t_u = timeit(lambda: unfused(False))
    t_us = timeit(lambda: unfused(True))
    fl = 4.0 * M * hidden * D
    print(f"M={M} hidden={hidden}: fused in place {t_f:.1f} us ({fl / t_f / 1e6:.0f} TFLOP/s)  fused (refetch x) {t_f0:.1f} us  "
          f"fused+save {t_fs:.1f} us  "
          f"unfused {t_u:.1f} us  unfused(train) {t_us:.1f} us", flush=True)
