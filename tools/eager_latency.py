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

from m3l_b200 import MAEExtractor
ext = MAEExtractor(None, mae, 256, False, 4).to(dev)
for use_graph in (False, True):
    ext.use_cuda_graph = use_graph
    mae.use_cuda_graph = use_graph
    for B in (64, 512):
        g = torch.Generator().manual_seed(B)
        obs = {"image": torch.rand(B, 4, 64, 64, 3, generator=g).to(dev), "tactile": (torch.rand(B, 4, 6, 32, 32, generator=g) * 2 - 1).to(dev)}
        w = torch.randn(B, 256, device=dev)
        xh, nh = bench.synth_batch(B, 7)
        x = {k: v.to(dev) for k, v in xh.items()}
        def it():
            ext.zero_grad(set_to_none=True)
            (ext(obs) * w).sum().backward()          # policy / value losses through the extractor
            mae(x).backward()                        # + the MAE reconstruction loss (ppo_mae.py:260-263,280)
        for _ in range(3): it()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): it()
        torch.cuda.synchronize()
        print(f"PPO-style minibatch B={B:4d} graphs={use_graph}: extractor fwd+bwd + mae fwd+bwd = {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")

# joint pass: one embedding of all tokens feeds the masked MAE pass and the extractor
ext.use_cuda_graph = True
mae.use_cuda_graph = True
for B in (64, 512):
    g = torch.Generator().manual_seed(B)
    obs = {"image": torch.rand(B, 4, 64, 64, 3, generator=g).to(dev), "tactile": (torch.rand(B, 4, 6, 32, 32, generator=g) * 2 - 1).to(dev)}
    w = torch.randn(B, 256, device=dev)
    def jt():
        ext.zero_grad(set_to_none=True)
        l = ext.joint_mae_loss(obs)
        ((ext(obs) * w).sum() + l).backward()
    for _ in range(3): jt()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): jt()
    torch.cuda.synchronize()
    print(f"PPO-style minibatch B={B:4d} JOINT pass (graphs): {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
