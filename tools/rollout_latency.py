"""Latency of MAEExtractor.forward (no grad) for small rollout batches (n_envs = 1 .. 64), device-resident obs."""
import sys, statistics, time
sys.path.insert(0, ".")
import torch
from m3l_b200 import VTT, VTMAE, MAEExtractor
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
torch.manual_seed(0)
enc = VTT(image_size=(64, 64), tactile_size=(32, 32), image_patch_size=8, tactile_patch_size=4, dim=256, depth=4, heads=4,
          mlp_dim=512, num_tactiles=2, image_channels=12, tactile_channels=12, frame_stack=4)
mae = VTMAE(encoder=enc, decoder_dim=256, masking_ratio=0.95, decoder_depth=3, decoder_heads=4, num_tactiles=2, frame_stack=4).to(dev)
ext = MAEExtractor(None, mae, 256, False, 4).to(dev)
for N in (1, 8, 64, 512):
    g = torch.Generator().manual_seed(N)
    obs = {"image": torch.rand(N, 4, 64, 64, 3, generator=g).to(dev), "tactile": (torch.rand(N, 4, 6, 32, 32, generator=g) * 2 - 1).to(dev)}
    with torch.no_grad():
        for _ in range(5): ext(obs)
        torch.cuda.synchronize()
        lat = []
        for _ in range(50):
            t0 = time.perf_counter(); y = ext(obs); torch.cuda.synchronize(); lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    print(f"n_envs {N:4d}: wall latency p50 {statistics.median(lat):.3f} ms  p90 {lat[44]:.3f} ms")
