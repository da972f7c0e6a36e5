"""The `torch.library` registration of the C-ABI entry points (namespace `m3l`, m3l_b200/torch_ops.py).

CPU: every op is registered with a schema and a fake (meta) implementation that yields the right shapes / dtypes
under FakeTensorMode, and has NO CPU kernel (no fallback).  GPU: `torch.library.opcheck` (schema, fake-tensor and
AOT-dispatch consistency against the real CUDA kernels) and values against plain torch fp32 statements."""
import pytest
import torch

import m3l_b200.torch_ops as T

OPS = torch.ops.m3l


def test_ops_are_registered_with_schemas():
    names = set(T.registered_ops())
    assert {"linear", "wgrad", "layernorm_fwd", "layernorm_bwd", "attention_fwd", "attention_bwd", "ln_mlp_fwd",
            "mask_indices", "patch_layernorm", "masked_patch_mse", "vt_load_image", "vt_load_tactile"} <= names
    for n in names:
        op = getattr(OPS, n).default
        assert op._schema.name == f"m3l::{n}"
        assert torch._C._dispatch_has_kernel_for_dispatch_key(f"m3l::{n}", "CUDA")
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(f"m3l::{n}", "CPU")          # no CPU fallback


def test_no_cpu_fallback_raises():
    with pytest.raises(NotImplementedError):
        OPS.linear(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(16, 64, dtype=torch.bfloat16))


def test_fake_implementations_give_shapes_and_dtypes():
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        bf = dict(dtype=torch.bfloat16, device="cuda")
        f32 = dict(dtype=torch.float32, device="cuda")
        a, w = torch.empty(1920, 256, **bf), torch.empty(768, 256, **bf)
        y = OPS.linear(a, w, torch.empty(768, **f32), None, 1, False)
        assert y.shape == (1920, 768) and y.dtype == torch.bfloat16 and y.device.type == "cuda"
        assert OPS.linear(a, w, None, None, 0, True).dtype == torch.float32
        g = OPS.wgrad(torch.empty(1920, 768, **bf), a)
        assert g.shape == (768, 256) and g.dtype == torch.float32
        xn, st = OPS.layernorm_fwd(a, torch.empty(256, **f32), torch.empty(256, **f32))
        assert xn.shape == a.shape and st.shape == (1920, 2) and st.dtype == torch.float32
        dx, dg, db = OPS.layernorm_bwd(a, a, st, torch.empty(256, **f32))
        assert dx.shape == a.shape and dg.shape == (256,) and db.dtype == torch.float32
        qkv = torch.empty(10 * 192, 768, **bf)
        o, lse = OPS.attention_fwd(qkv, 10, 192, 4, 64, 0.125)
        assert o.shape == (1920, 256) and lse.shape == (10, 4, 192)
        assert OPS.attention_bwd(qkv, o, o, lse, 10, 192, 4, 64, 0.125).shape == qkv.shape
        out = OPS.ln_mlp_fwd(a, torch.empty(256, **f32), torch.empty(256, **f32), torch.empty(1024, 256, **bf),
                             torch.empty(1024, **f32), torch.empty(256, 1024, **bf), torch.empty(256, **f32))
        assert out.shape == a.shape and out.dtype == torch.bfloat16
        m, u = OPS.mask_indices(torch.empty(10, 192, **f32), [0, 64, 128], [64, 64, 64], [60, 61, 61])
        assert m.shape == (10, 182) and u.shape == (10, 10) and m.dtype == torch.int64
        img = torch.empty(10, 12, 64, 64, **f32)
        pl = OPS.patch_layernorm([img], 8, 8, 0, u, 0, 4, torch.empty(768, **f32), torch.empty(768, **f32))
        assert pl.shape == (40, 768) and pl.dtype == torch.bfloat16
        loss, dp = OPS.masked_patch_mse([img], 8, 8, 0, m, 0, 60, torch.empty(600, 768, **f32), 1.0)
        assert loss.shape == () and dp.shape == (600, 768) and dp.dtype == torch.bfloat16


# ------------------------------------------------------------------------------------------------ GPU
DEV = "cuda"


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))


@pytest.mark.gpu
def test_opcheck_and_values_linear_layernorm_mlp():
    torch.manual_seed(0)
    a = torch.randn(384, 256, device=DEV).bfloat16()
    w = (torch.randn(512, 256, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(512, device=DEV)
    torch.library.opcheck(OPS.linear.default, (a, w, bias, None, 0, False))
    torch.library.opcheck(OPS.linear.default, (a, w, None, None, 0, True))
    y = OPS.linear(a, w, bias)
    assert _cos(y, a.float() @ w.float().T + bias) > 0.9999
    dy = torch.randn(384, 512, device=DEV).bfloat16()
    torch.library.opcheck(OPS.wgrad.default, (dy, a))
    assert _cos(OPS.wgrad(dy, a), dy.float().T @ a.float()) > 0.9999
    g, b = torch.rand(256, device=DEV) + 0.5, torch.randn(256, device=DEV)
    torch.library.opcheck(OPS.layernorm_fwd.default, (a, g, b))
    xn, st = OPS.layernorm_fwd(a, g, b)
    assert _cos(xn, torch.nn.functional.layer_norm(a.float(), (256,), g, b)) > 0.9999
    torch.library.opcheck(OPS.layernorm_bwd.default, (a, a, st, g))
    w1 = (torch.randn(1024, 256, device=DEV) * 0.05).bfloat16(); b1 = torch.randn(1024, device=DEV) * 0.1
    w2 = (torch.randn(256, 1024, device=DEV) * 0.05).bfloat16(); b2 = torch.randn(256, device=DEV) * 0.1
    torch.library.opcheck(OPS.ln_mlp_fwd.default, (a, g, b, w1, b1, w2, b2))
    out = OPS.ln_mlp_fwd(a, g, b, w1, b1, w2, b2)
    xr = torch.nn.functional.layer_norm(a.float(), (256,), g, b)
    ref = a.float() + torch.nn.functional.gelu(xr @ w1.float().T + b1) @ w2.float().T + b2
    assert _cos(out, ref) > 0.9995


@pytest.mark.gpu
def test_opcheck_attention_mask_patch_mse():
    torch.manual_seed(1)
    B, n, H, dh = 3, 192, 4, 64
    qkv = (torch.randn(B * n, 3 * H * dh, device=DEV) * 0.5).bfloat16()
    torch.library.opcheck(OPS.attention_fwd.default, (qkv, B, n, H, dh, dh ** -0.5))
    o, lse = OPS.attention_fwd(qkv, B, n, H, dh, dh ** -0.5)
    q, k, v = [t.view(B, n, H, dh).transpose(1, 2).float() for t in qkv.chunk(3, -1)]
    ref = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1) @ v
    assert _cos(o, ref.transpose(1, 2).reshape(B * n, H * dh)) > 0.9995
    torch.library.opcheck(OPS.attention_bwd.default, (qkv, o, o, lse, B, n, H, dh, dh ** -0.5))
    noise = torch.rand(5, 192, device=DEV)
    torch.library.opcheck(OPS.mask_indices.default, (noise, [0, 64, 128], [64, 64, 64], [60, 61, 61]))
    m, u = OPS.mask_indices(noise, [0, 64, 128], [64, 64, 64], [60, 61, 61])
    perm = noise[:, :64].argsort(dim=1, stable=True)
    assert torch.equal(m[:, :60], perm[:, :60]) and torch.equal(u[:, :4], perm[:, 60:])
    img = torch.rand(5, 12, 64, 64, device=DEV)
    g, b = torch.ones(768, device=DEV), torch.zeros(768, device=DEV)
    torch.library.opcheck(OPS.patch_layernorm.default, ([img], 8, 8, 0, u, 0, 4, g, b))
    pred = torch.randn(5 * 60, 768, device=DEV)
    torch.library.opcheck(OPS.masked_patch_mse.default, ([img], 8, 8, 0, m, 0, 60, pred, 1.0))
    loss, dp = OPS.masked_patch_mse([img], 8, 8, 0, m, 0, 60, pred, 1.0)
    patches = img.view(5, 12, 8, 8, 8, 8).permute(0, 2, 4, 3, 5, 1).reshape(5, 64, 768)
    tgt = patches[torch.arange(5, device=DEV)[:, None], m[:, :60]].reshape(300, 768)
    assert abs(float(loss) - float(torch.nn.functional.mse_loss(pred, tgt))) < 1e-4 * float(loss)
    obs_i = torch.rand(2, 4, 64, 64, 3, device=DEV)
    obs_t = torch.rand(2, 4, 6, 32, 32, device=DEV) * 2 - 1
    torch.library.opcheck(OPS.vt_load_image.default, (obs_i, 4))
    torch.library.opcheck(OPS.vt_load_tactile.default, (obs_t, 4, 1))
    assert torch.equal(OPS.vt_load_image(obs_i, 4), obs_i.permute(0, 1, 4, 2, 3).reshape(2, 12, 64, 64))
    assert torch.equal(OPS.vt_load_tactile(obs_t, 4, 1), (obs_t[:, :, 3:6].reshape(2, 12, 32, 32) + 1) / 2)
