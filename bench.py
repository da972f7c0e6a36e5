#!/usr/bin/env python
"""Benchmark of the VTMAE train step (BASELINE.json metric: train samples/sec, fwd+bwd+AdamW).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the canonical train.py model (64x64x12 image frame-stack +
two 32x32x12 tactile maps, dim 256, enc 4x / dec 3x, mask 0.95, early_conv_masking=False),
batch 256 per GPU, bf16 tensor-core compute with fp32 master weights / accumulation, synthetic
data, random-init weights.  One "step" = zero_grad + forward + backward (+ gradient all-reduce for
N > 1) + clip_grad_norm_(0.5) + AdamW, i.e. /root/reference/models/pretrain_models.py:707-711.

Prints ONE JSON line on rank 0.  `value` = whole-job samples/s with inputs resident in HBM;
`e2e` = the same metric through the public module API with pinned HOST inputs (H2D copy of every
step's batch and a D2H read of the loss inside the timed region).  `--impl reference` times the
reference algorithm's CPU path (the oracle port of the reference's PyTorch code; the reference tree
itself does not exist on the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOPS_PER_SAMPLE = 3404.7e6  # fwd + bwd, SURVEY.md §8(d) / BASELINE.md §3 (sum of 2*M*N*K, x3)
BATCH_PER_GPU = 256
WORKLOAD = "VTMAE pretrain step, canonical train.py model, batch 256/GPU, bf16 (BASELINE.json configs[1])"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
                self.lines = [l for l in out.splitlines() if l.strip()]
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_model(device):
    import torch
    from m3l_b200 import VTT, VTMAE
    torch.manual_seed(0)
    enc = VTT(image_size=(64, 64), tactile_size=(32, 32), image_patch_size=8, tactile_patch_size=4, dim=256, depth=4,
              heads=4, mlp_dim=512, num_tactiles=2, image_channels=12, tactile_channels=12, frame_stack=4)
    mae = VTMAE(encoder=enc, decoder_dim=256, masking_ratio=0.95, decoder_depth=3, decoder_heads=4, num_tactiles=2,
                early_conv_masking=False, frame_stack=4)
    return mae.to(device)


def synth_batch(B, seed, pinned=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = {"image": torch.rand(B, 12, 64, 64, generator=g), "tactile1": torch.rand(B, 12, 32, 32, generator=g),
         "tactile2": torch.rand(B, 12, 32, 32, generator=g)}
    noise = torch.rand(B, 192, generator=g)
    if pinned:
        x = {k: v.pin_memory() for k, v in x.items()}
        noise = noise.pin_memory()
    return x, noise


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, batch: int = 32):
    """Times zero_grad+fwd+bwd+clip+AdamW of the oracle (port of the reference's PyTorch code) in fp32
    on all host cores, on a `batch`-sample slice of the workload.  Returns (samples/s, ms/step, cores)."""
    import torch
    from oracle import vtmae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.VTMAEConfig()
    sd = O.init_state_dict(cfg, seed=0)
    x, _ = synth_batch(batch, 1234)
    g = torch.Generator().manual_seed(1)
    noise = O.tie_free_noise(batch, 192, g, [64, 64, 64])
    st = O.AdamWState()
    for _ in range(warmup):
        O.train_step(sd, cfg, x, noise, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(sd, cfg, x, noise, st)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    sps, ms, cores = cpu_reference_run(steps, max(1, min(args.warmup, 3)))
    sample = f"batch 32 slice of the 256/GPU workload, fp32, torch CPU, {cores} threads, {steps} timed steps"
    line = {"impl": "reference", "metric": "VTMAE train samples/sec (fwd+bwd+AdamW)", "value": sps, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=10):
    import torch
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3  # seconds


def dominant_kernel_roofline(peaks):
    """Live CUDA-event timing of the dominant kernel (the tcgen05 GEMM on the decoder feed-forward
    shape M=49152, N=1024, K=256: 25.8 GFLOP per launch) on the current stream."""
    import torch
    from m3l_b200 import ops
    M, N, K = BATCH_PER_GPU * 192, 1024, 256
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    sec = time_kernel(lambda: ops.gemm(a, b, out=out, bias=bias, act=ops.GELU_FWD, aux_out=pre))
    flops = 2.0 * M * N * K
    # algorithmic HBM bytes per launch: A (M x K bf16) + W (N x K bf16) + bias + the two bf16 outputs
    # (GELU(pre) and GELU'(pre), both M x N) = 227.0 MB -> 113.8 flop/B, below the ridge (burst 1684 TF/s /
    # 6460 GB/s = 261 flop/B): the kernel's binding roofline is HBM; the tensor-pipe view is kept beside it.
    bytes_alg = 2.0 * M * K + 2.0 * N * K + 4.0 * N + 2 * 2.0 * M * N
    gbs = bytes_alg / sec / 1e9
    return {"bound": "hbm", "kernel": "gemm_gelu16_kernel decoder FF1 (M=49152,N=1024,K=256,+bias, writes GELU and GELU')",
            "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full capture
            # profiles/r01_ncu_full_gemm_v4.txt (gemm_gelu16_kernel: 25.72 MB read + 147.36 MB written; the rest of the
            # 201 MB of outputs was still dirty in the 126 MB L2 when the kernel ended)
            "traffic": 173.1e6, "algorithmic_bytes": bytes_alg, "us_per_launch": sec * 1e6,
            "peak_source": peaks["source"] + ", burst",
            "tensor_view": {"achieved": flops / sec / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                            "frac": flops / sec / 1e12 / peaks["tf_burst"], "flop_per_byte": flops / bytes_alg}}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from m3l_b200.trainer import FusedTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B = BATCH_PER_GPU
    mae = build_model(dev)
    mae._sync()
    trainer = FusedTrainer(mae, lr=1e-4, use_cuda_graph=not args.no_graph)
    mae._trainer = trainer

    # rotating resident batches: 4 x 75.5 MB inputs; together with ~2 GB of activations written per step
    # the working set is far larger than the 126 MB L2
    nbuf = 4
    host = [synth_batch(B, 1234 + rank * 100 + i, pinned=True) for i in range(nbuf)]
    devb = [({k: v.to(dev) for k, v in x.items()}, n.to(dev)) for x, n in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        x, n = devb[i % nbuf]
        trainer.step(x, noise=n)
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        s.record()
        for i in range(args.steps):
            x, n = devb[i % nbuf]
            loss = trainer.step(x, noise=n)
        e.record()
        barrier()
    ms = s.elapsed_time(e)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)
    loss_val = float(loss.item())
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_per_step, "value": value, "loss": loss_val}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end: pinned host inputs, H2D every step (prefetched one step ahead on a copy
    #      stream), D2H read of the loss every step ------------------------------------------
    copy_stream = torch.cuda.Stream()
    stage = [({k: torch.empty_like(v, device=dev) for k, v in host[0][0].items()}, torch.empty(B, 192, device=dev))
             for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        slot = i % 2
        x, n = host[i % nbuf]
        copy_stream.wait_event(consumed[slot])
        with torch.cuda.stream(copy_stream):
            for k, v in x.items():
                stage[slot][0][k].copy_(v, non_blocking=True)
            stage[slot][1].copy_(n, non_blocking=True)
            ready[slot].record(copy_stream)

    for ev in consumed:
        ev.record()
    h2d = sum(v.numel() * v.element_size() for v in host[0][0].values()) + host[0][1].numel() * 4
    e2e_steps = args.steps
    # H2D alone (reported next to e2e: the floor the PCIe link sets for a step)
    barrier()
    hs, he = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hs.record(copy_stream)
    for i in range(4):
        prefetch(i)
        consumed[i % 2].record(copy_stream)
    he.record(copy_stream)
    barrier()
    h2d_ms = hs.elapsed_time(he) / 4
    for ev in consumed:
        ev.record()
    # every step's loss is copied D2H into pinned memory on the compute stream and READ by the host one
    # step later (after its event), so the host never drains the GPU queue; all K losses are read
    # inside the timed region
    loss_host = torch.empty(e2e_steps, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(e2e_steps)]
    barrier()
    t0 = time.perf_counter()
    s.record()
    prefetch(0)
    losses = []
    for i in range(e2e_steps):
        slot = i % 2
        if i + 1 < e2e_steps:
            prefetch(i + 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        l = trainer.step(stage[slot][0], noise=stage[slot][1])
        consumed[slot].record()
        loss_host[i:i + 1].copy_(l.reshape(1), non_blocking=True)     # D2H of the step result
        loss_ev[i].record()
        if i > 0:
            loss_ev[i - 1].synchronize()
            losses.append(float(loss_host[i - 1]))
    loss_ev[e2e_steps - 1].synchronize()
    losses.append(float(loss_host[e2e_steps - 1]))
    e.record()
    barrier()
    assert len(losses) == e2e_steps and all(x == x for x in losses)
    ms_e2e = s.elapsed_time(e)
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / (float(t.item()) * 1e-3)
    wall_e2e = time.perf_counter() - t0

    if rank == 0:
        roof = dominant_kernel_roofline(peaks)
        step_tflops = value / world * FLOPS_PER_SAMPLE / 1e12
        # CPU baseline: rank 0 at N = 1 only (torchrun pins OMP_NUM_THREADS=1; the driver's reference arm times it)
        cpu_sps, cpu_ms, cores = cpu_reference_run(steps=30, warmup=3) if world == 1 else (None, None, None)
        launches = (trainer.kernel_launches_per_step or 0)
        line = {
            "metric": "VTMAE train samples/sec (fwd+bwd+AdamW)", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "batch_per_gpu": B, "parallelism": f"dp{world}",
                       "l2": "inputs rotate over 4 resident batches (302 MB) and every step streams >2 GB of "
                             "activations through HBM, far above the 126 MB L2; no explicit flush",
                       "cuda_graph": bool(trainer.use_graph), "loss_last_step": loss_val},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "note": "pinned host batch -> H2D prefetched one step ahead on a copy stream; every step's loss "
                            "copied D2H to pinned memory and read by the host one step later",
                    "h2d_ms_per_step_alone": h2d_ms, "wall_s": wall_e2e},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "roofline": roof,
            "step_tensor_utilisation": {"achieved": step_tflops, "unit": "TFLOP/s (3404.7 MFLOP/sample x samples/s/GPU)",
                                        "peak": peaks["tf_sustained"], "frac": step_tflops / peaks["tf_sustained"],
                                        "peak_source": peaks["source"] + ", sustained"},
        }
        if cpu_sps is not None:
            line["cpu_baseline"] = {"value": cpu_sps, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": f"batch 32 slice of the workload, fp32 torch CPU oracle, 30 timed steps, "
                                              f"{cpu_ms:.1f} ms/step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="m3l_b200", choices=["m3l_b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly (for ncu launch lists)")
    ap.add_argument("--profile", action="store_true", help="device-resident loop only (no e2e / CPU baseline legs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
