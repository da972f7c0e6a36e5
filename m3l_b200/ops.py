"""Python wrappers of the individual C-ABI kernels (unit-test and composition surface).

Every function takes CUDA torch tensors, passes raw pointers + the current stream to the
C-ABI and returns torch tensors.  torch is only the allocator / stream provider here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs, check, current_stream, ptr

GELU_FWD = 1
GELU_BWD = 2


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.M3LError("m3l_b200 ops need CUDA tensors (there is no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, *, mn_major: bool = False, out: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.bfloat16, accumulate: bool = False, splits: int = 1, bn: int = 0,
         bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act: int = 0,
         aux_out: Optional[torch.Tensor] = None, aux_in: Optional[torch.Tensor] = None,
         alpha: float = 1.0) -> torch.Tensor:
    """out[m, n] = epilogue(alpha * sum_k A[m, k] B[n, k]).

    mn_major=False: a is [M, K], b is [N, K] (nn.Linear forward: x @ W.T).
    mn_major=True : a is [K, M], b is [K, N] (wgrad: a.T @ b), optionally split-K + accumulate.
    """
    _req_cuda(a, b, out, bias, residual, aux_out, aux_in)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if mn_major:
        K, M = a.shape
        K2, N = b.shape
    else:
        M, K = a.shape
        N, K2 = b.shape
    assert K == K2, (a.shape, b.shape)
    if out is None:
        dt = torch.float32 if accumulate else out_dtype
        out = (torch.zeros if accumulate else torch.empty)((M, N), dtype=dt, device=a.device)
    assert out.stride(1) == 1
    if accumulate:
        assert out.dtype == torch.float32
        out_mode = 2
    else:
        out_mode = 0 if out.dtype == torch.bfloat16 else 1
    args = GemmArgs()
    args.a, args.b = ptr(a), ptr(b)
    args.lda, args.ldb = a.stride(0), b.stride(0)
    args.a_mn_major = args.b_mn_major = 1 if mn_major else 0
    args.m, args.n, args.k = M, N, K
    args.splits, args.bn = splits, bn
    args.out, args.ldo, args.out_mode = ptr(out), out.stride(0), out_mode
    args.bias = ptr(bias)
    args.residual, args.ldr = ptr(residual), (residual.stride(0) if residual is not None else 0)
    args.act = act
    args.aux_out, args.aux_in = ptr(aux_out), ptr(aux_in)
    aux = aux_out if aux_out is not None else aux_in
    args.ld_aux = aux.stride(0) if aux is not None else 0
    args.alpha = alpha
    check(_lib.load().m3l_gemm_bf16(C.byref(args), current_stream()), "m3l_gemm_bf16")
    return out
