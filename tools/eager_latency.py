"""Wall time of the drop-in autograd path  loss = mae(x); loss.backward()  (what PPO_MAE / SAC_MAE call) vs the fused
CUDA-graph train step, per batch size."""
import sys, time
sys.path.insert(0, ".")
import torch
import bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
mae = bench.build_model(dev)
mae._sync()
for B in (32, 256):
    xh, nh = bench.synth_batch(B, 7)
    x = {k: v.to(dev) for k, v in xh.items()}; n = nh.to(dev)
    for _ in range(3):
        mae.zero_grad(set_to_none=True); l = mae(x, noise=n); l.backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        mae.zero_grad(set_to_none=True); l = mae(x, noise=n); l.backward()
    torch.cuda.synchronize()
    print(f"B={B:4d} eager autograd fwd+bwd: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
