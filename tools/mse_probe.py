"""CUDA-event timings of the masked-patch MSE kernel at the benchmark shapes (image head, tactile head),
with and without the fused bias-gradient column sums; L2 flushed between launches."""
import sys
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, it=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / it * 1e3


B = 256
img = torch.rand(B, 12, 64, 64, device=dev)
tac = [torch.rand(B, 12, 32, 32, device=dev) for _ in range(2)]
for name, maps, pp, ntok, nm in (("image", [img], 8, 64, 60), ("tactile", tac, 4, 128, 122)):
    ps = ops.make_patch_source(maps, pp, pp, 0)
    P = pp * pp * 12
    idx = torch.stack([torch.randperm(ntok)[:nm] for _ in range(B)]).to(dev)
    pred = torch.randn(B * nm, P, device=dev)
    acc = torch.zeros(1, device=dev)
    dpred = torch.empty(B * nm, P, device=dev, dtype=torch.bfloat16)
    dcs = torch.zeros(P, device=dev)
    mb = B * nm * P * 10 / 1e6
    a = t(lambda: ops.mse_loss(ps, B, nm, pred, 1e-6, acc, tok_idx=idx, dpred=dpred, dpred_colsum=dcs))
    b = t(lambda: ops.mse_loss(ps, B, nm, pred, 1e-6, acc, tok_idx=idx, dpred=dpred))
    print(f"mse {name:8s} rows {B * nm:6d} P {P:4d}  {mb:6.1f} MB   with colsum {a:6.1f} us ({mb / a * 1e3:5.0f} GB/s)   without {b:6.1f} us")
