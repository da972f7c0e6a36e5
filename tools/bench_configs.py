"""Secondary workloads of BASELINE.json on ONE GPU (the bench.py line stays configs[1]):
  cfg3  vision_only_control VTMAE train step (num_tactiles=0), 128 and 1024 samples / GPU
  cfg4  DINO-tac-MAE, MAE side: tactile-only 70x70 / patch 14 / dim 384 train step, batch 512
  cfg5  rollout-time MAEExtractor.forward (no masking, no grad) on 4096 env observations: latency + obs/s,
        device-resident observations and host (pinned) observations incl. H2D
Prints one JSON object per workload.  usage: python tools/bench_configs.py [cfg3 cfg4 cfg5]"""
import json
import statistics
import sys

sys.path.insert(0, ".")
import torch

from m3l_b200 import VTT, VTMAE, MAEExtractor
from m3l_b200.trainer import FusedTrainer

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]


def time_steps(fn, warm=5, iters=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def train_bench(name, mae, batches, flops_per_sample):
    mae._sync()
    tr = FusedTrainer(mae, lr=1e-4)
    k = [0]

    def step():
        x, n = batches[k[0] % len(batches)]
        k[0] += 1
        tr.step(x, noise=n)

    ms = time_steps(step)
    B = batches[0][1].shape[0]
    sps = B / (ms * 1e-3)
    print(json.dumps({"workload": name, "batch": B, "ms_per_step": ms, "samples_per_s": sps,
                      "tflops": sps * flops_per_sample / 1e12, "launches_per_step": tr.kernel_launches_per_step}), flush=True)


if "cfg3" in which:
    for B in (128, 1024):
        torch.manual_seed(0)
        enc = VTT(image_size=(64, 64), tactile_size=(32, 32), image_patch_size=8, tactile_patch_size=4, dim=256, depth=4,
                  heads=4, mlp_dim=512, num_tactiles=0, image_channels=12, tactile_channels=12, frame_stack=4)
        mae = VTMAE(encoder=enc, decoder_dim=256, masking_ratio=0.95, decoder_depth=3, decoder_heads=4, num_tactiles=0,
                    frame_stack=4).to(dev)
        g = torch.Generator().manual_seed(1)
        batches = [({"image": torch.rand(B, 12, 64, 64, generator=g).to(dev)}, torch.rand(B, 64, generator=g).to(dev))
                   for _ in range(3)]
        train_bench(f"cfg3 vision_only_control VTMAE train step, {B}/GPU", mae, batches, 1140.5e6)
        del mae, batches

if "cfg4" in which:
    B = 512
    torch.manual_seed(0)
    enc = VTT(image_size=(70, 70), tactile_size=(70, 70), image_patch_size=14, tactile_patch_size=14, dim=384, depth=4,
              heads=4, mlp_dim=768, num_tactiles=2, image_channels=12, tactile_channels=12, frame_stack=4)
    mae = VTMAE(encoder=enc, decoder_dim=384, masking_ratio=0.8, decoder_depth=3, decoder_heads=4, num_tactiles=2,
                frame_stack=4).to(dev)
    g = torch.Generator().manual_seed(2)
    batches = [({f"tactile{i + 1}": torch.rand(B, 12, 70, 70, generator=g).to(dev) for i in range(2)},
                torch.rand(B, 50, generator=g).to(dev)) for _ in range(3)]
    train_bench("cfg4 DINO-tac-MAE (MAE side, tactile-only, dim 384) train step, 512/GPU", mae, batches, 2163.5e6)
    del mae, batches

if "cfg5" in which:
    N = 4096
    torch.manual_seed(0)
    enc = VTT(image_size=(64, 64), tactile_size=(32, 32), image_patch_size=8, tactile_patch_size=4, dim=256, depth=4,
              heads=4, mlp_dim=512, num_tactiles=2, image_channels=12, tactile_channels=12, frame_stack=4)
    mae = VTMAE(encoder=enc, decoder_dim=256, masking_ratio=0.95, decoder_depth=3, decoder_heads=4, num_tactiles=2,
                frame_stack=4).to(dev)
    ext = MAEExtractor(None, mae, 256, False, 4).to(dev)
    g = torch.Generator().manual_seed(3)
    obs_h = {"image": torch.rand(N, 4, 64, 64, 3, generator=g).pin_memory(),
             "tactile": (torch.rand(N, 4, 6, 32, 32, generator=g) * 2 - 1).pin_memory()}
    obs_d = {k: v.to(dev) for k, v in obs_h.items()}
    for label, obs in (("device-resident obs", obs_d), ("pinned host obs (H2D inside)", obs_h)):
        lat = []
        with torch.no_grad():
            for _ in range(5):
                ext(obs)
            torch.cuda.synchronize()
            for _ in range(100):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                y = ext(obs)
                e.record()
                torch.cuda.synchronize()
                lat.append(s.elapsed_time(e))
        lat.sort()
        p50, p99 = statistics.median(lat), lat[98]
        print(json.dumps({"workload": f"cfg5 rollout MAEExtractor.forward, {N} obs, {label}", "out_shape": list(y.shape),
                          "latency_ms_p50": p50, "latency_ms_p99": p99, "obs_per_s": N / (p50 * 1e-3),
                          "tflops": N / (p50 * 1e-3) * 1233.1e6 / 1e12}), flush=True)
