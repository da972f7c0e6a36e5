"""Where the step time goes INSIDE CUDA graphs (PDL overlap included): each phase of the fused step is
captured as its own graph and timed with CUDA events.  usage: python tools/step_breakdown.py [B]"""
import sys
sys.path.insert(0, ".")
import torch
import bench
from m3l_b200 import engine, ops
from m3l_b200.trainer import FusedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
model = bench.build_model(dev)
A = model._sync()
tr = FusedTrainer(model, use_cuda_graph=False)
x_host, noise_host = bench.synth_batch(B, 1234)
xs = {k: v.to(dev) for k, v in x_host.items()}
noise = noise_host.to(dev)
xs, geo, _ = model._prep_inputs(xs, True, True)
live, ranges, _ = tr._plan(geo)
G = engine.GradView(A, tr.gflat)


def graph_time(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    cnt = ops.LaunchCounter()
    with cnt, torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, cnt.count


rows = []
box = {}


def full():
    tr._phase_a(xs, noise, geo, box)
    tr._phase_b(box)
    tr._phase_c(ranges)


rows.append(("full step", *graph_time(full)))


def fwd_only():
    box["l"], box["c"] = engine.mae_forward(model, xs, noise, geo, training=True, gflat=tr.gflat)


rows.append(("mae_forward (training)", *graph_time(fwd_only)))
ctx = box["c"]
rows.append(("backward: heads + decoder", *graph_time(lambda: engine.mae_backward_decoder(model, ctx, tr.gflat))))
rows.append(("backward: encoder + embed", *graph_time(lambda: engine.mae_backward_encoder(model, ctx, tr.gflat))))
rows.append(("optimizer (sumsq, clip+AdamW, shadows)", *graph_time(lambda: tr._phase_c(ranges))))

for name, spec, n in (("encoder", model.enc_spec, geo.nv), ("decoder", model.dec_spec, geo.n)):
    x0 = torch.randn(B * n, spec.dim, device=dev).bfloat16()
    saved = []
    engine.stack_fwd(A, spec, x0, B, n, saved)

    def f(spec=spec, x0=x0, n=n):
        engine.stack_fwd(A, spec, x0, B, n, [])

    def b(spec=spec, x0=x0, n=n, saved=saved):
        engine.stack_bwd(A, G, spec, x0, B, n, saved)

    rows.append((f"{name} stack fwd (n={n}, depth {spec.depth})", *graph_time(f)))
    rows.append((f"{name} stack bwd", *graph_time(b)))

for name, us, k in rows:
    print(f"{name:44s} {us:9.1f} us  {k:4d} launches")
