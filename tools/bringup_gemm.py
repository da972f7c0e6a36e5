"""GPU bring-up of the tcgen05 GEMM: many configs, reports errors instead of stopping at the first."""
import sys, time, traceback
sys.path.insert(0, ".")
import torch
from m3l_b200 import ops

torch.manual_seed(0)
dev = "cuda"

def ref_err(out, ref):
    out = out.float(); ref = ref.float()
    return ((out - ref).abs().max() / (ref.abs().max() + 1e-9)).item()

def run(name, fn):
    try:
        e = fn()
        torch.cuda.synchronize()
        print(f"{name}: rel_max_err={e:.3e} {'OK' if e < 2e-2 else 'BAD'}", flush=True)
    except Exception as ex:
        print(f"{name}: EXC {type(ex).__name__}: {ex}", flush=True)
        traceback.print_exc()

def kk(M, N, K, bn, **kw):
    def f():
        a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
        out = ops.gemm(a, b, bn=bn, **kw)
        return ref_err(out, a.float() @ b.float().T)
    return f

def mm(M, N, K, bn, splits):
    def f():
        a = torch.randn(K, M, device=dev).bfloat16(); b = torch.randn(K, N, device=dev).bfloat16()
        out = ops.gemm(a, b, mn_major=True, bn=bn, splits=splits, accumulate=True)
        return ref_err(out, a.float().T @ b.float())
    return f

for bn in (64, 128, 256):
    run(f"KK 128x{bn}x64 bn{bn}", kk(128, bn, 64, bn))
    run(f"KK 256x512x256 bn{bn}", kk(256, 512, 256, bn))
    run(f"KK ragged 1000x520x200 bn{bn}", kk(1000, 520, 200, bn))
    run(f"KK big 49152x768x256 bn{bn}", kk(49152, 768, 256, bn))
    run(f"MM 128x{bn}x64 bn{bn}", mm(128, bn, 64, bn, 1))
    run(f"MM 256x768x4096 s4 bn{bn}", mm(256, 768, 4096, bn, 4))
    run(f"MM ragged 264x200x1000 s3 bn{bn}", mm(264, 200, 1000, bn, 3))

# epilogue variants
def epi():
    M, N, K = 512, 256, 256
    a = torch.randn(M, K, device=dev).bfloat16(); b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev).bfloat16()
    out = ops.gemm(a, b, bias=bias, residual=res)
    ref = a.float() @ b.float().T + bias + res.float()
    e1 = ref_err(out, ref)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    h = ops.gemm(a, b, bias=bias, act=ops.GELU_FWD, aux_out=pre)
    z = a.float() @ b.float().T + bias
    e2 = max(ref_err(h, torch.nn.functional.gelu(z)), ref_err(pre, z))
    g = ops.gemm(a, b, act=ops.GELU_BWD, aux_in=pre)
    zz = pre.float().requires_grad_(True); torch.nn.functional.gelu(zz).sum().backward()
    e3 = ref_err(g, (a.float() @ b.float().T) * zz.grad)
    print(f"  epi errs: bias+res {e1:.2e} gelu {e2:.2e} gelu' {e3:.2e}")
    return max(e1, e2, e3)
run("epilogues", epi)

# timing of the big decoder shapes
def bench(M, N, K, bn, iters=20):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, b, out=out, bn=bn)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): ops.gemm(a, b, out=out, bn=bn)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    tf = 2 * M * N * K / ms / 1e9
    s.record()
    for _ in range(iters): torch.matmul(a, b.T, out=out)
    e.record(); torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    print(f"bench {M}x{N}x{K} bn{bn}: {ms*1e3:.1f} us {tf:.0f} TFLOP/s | cublas {ms2*1e3:.1f} us {2*M*N*K/ms2/1e9:.0f} TFLOP/s", flush=True)
for (M, N, K) in [(49152, 768, 256), (49152, 256, 256), (49152, 1024, 256), (49152, 256, 1024), (8192, 8192, 8192)]:
    for bn in (128, 256):
        try: bench(M, N, K, bn)
        except Exception as ex: print("bench EXC", ex)
def bench_w(M, N, K, bn, splits, iters=20):
    a = torch.randn(K, M, device=dev).bfloat16(); b = torch.randn(K, N, device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev)
    for _ in range(3): ops.gemm(a, b, mn_major=True, out=out, accumulate=True, splits=splits, bn=bn)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): ops.gemm(a, b, mn_major=True, out=out, accumulate=True, splits=splits, bn=bn)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"wgrad {M}x{N}x{K} bn{bn} s{splits}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
for (M, N, K, bn, s) in [(1024, 256, 49152, 128, 9), (1024, 256, 49152, 256, 37), (768, 256, 49152, 128, 12), (256, 256, 49152, 128, 37), (256, 1024, 49152, 256, 37)]:
    try: bench_w(M, N, K, bn, s)
    except Exception as ex: print("bench_w EXC", ex)
