"""GPU parity of m3l_b200.vtt.VTT (drop-in for /root/reference/models/VTT.py::VTT) against the CPU
oracle (oracle/vtt_dino_oracle.py, pinned bit-for-bit against the unmodified reference in
tests/test_oracle_vs_reference.py).  Tolerances as for the MAE path: outputs cosine >= 0.9995, every parameter
gradient cosine >= 0.999 (bf16 tensor-core compute, fp32 LayerNorm / softmax / accumulation)."""
import pytest
import torch

from oracle import vtt_dino_oracle as VD

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cos(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def build(cfg, seed):
    from m3l_b200.vtt import VTT
    torch.manual_seed(seed)
    m = VTT(image_size=cfg.image_size, tactile_size=cfg.tactile_size, image_patch_size=cfg.image_patch_size,
            tactile_patch_size=cfg.tactile_patch_size, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim,
            num_tactiles=cfg.num_tactiles, image_channels=cfg.image_channels, tactile_channels=cfg.tactile_channels,
            dim_head=cfg.dim_head, num_register_tokens=cfg.num_register_tokens, pos_embed_fn="sinusoidal")
    with torch.no_grad():
        if m.register_tokens is not None:
            m.register_tokens.normal_(0, 0.5)
        for k, p in m.named_parameters():       # timm init has zero biases / unit LN: perturb so every term matters
            if k.endswith(".bias"):
                p.normal_(0, 0.05)
            elif p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    return m


@pytest.mark.parametrize("regs,n_masks,keep,heads", [(1, 0, 0, 8), (1, 2, 40, 8), (0, 1, 17, 4), (2, 3, 8, 4)])
def test_forward_features_and_grads_vs_oracle(regs, n_masks, keep, heads):
    cfg = VD.VTTDinoConfig(depth=2, num_register_tokens=regs, heads=heads)
    m = build(cfg, seed=20 + regs)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    g = torch.Generator().manual_seed(31)
    B = 3
    x = {"image": torch.rand(B, 12, 64, 64, generator=g), "tactile1": torch.rand(B, 12, 32, 32, generator=g),
         "tactile2": torch.rand(B, 12, 32, 32, generator=g)}
    masks = [torch.stack([torch.randperm(64, generator=g)[:keep] for _ in range(B)]) for _ in range(n_masks)] or None
    out = m.forward_features({k: v.to(DEV) for k, v in x.items()}, [mk.to(DEV) for mk in masks] if masks else None)
    for k in sd:
        if sd[k].dtype.is_floating_point and k != "pos_embed.frequency_bands":
            sd[k].requires_grad_(True)
    ref = VD.forward_features(sd, cfg, x, masks)
    for k in ("x_norm_regtokens", "x_norm_patchtokens", "x_prenorm"):
        assert out[k].shape == ref[k].shape and out[k].dtype == torch.float32, k
        if out[k].numel():
            assert cos(out[k], ref[k]) >= 0.9995, (k, cos(out[k], ref[k]))
    w1 = torch.randn(ref["x_norm_patchtokens"].shape, generator=g)
    w2 = torch.randn(ref["x_prenorm"].shape, generator=g)
    ((out["x_norm_patchtokens"] * w1.to(DEV)).sum() + (out["x_norm_regtokens"] ** 2).sum()
     + 0.5 * (out["x_prenorm"] * w2.to(DEV)).sum()).backward()
    ((ref["x_norm_patchtokens"] * w1).sum() + (ref["x_norm_regtokens"] ** 2).sum() + 0.5 * (ref["x_prenorm"] * w2).sum()).backward()
    named = dict(m.named_parameters())
    for k, p in named.items():
        gr = sd[k].grad
        if gr is None or float(gr.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            assert p.grad is not None and cos(p.grad, gr) >= 0.999, (k, cos(p.grad, gr) if p.grad is not None else None)


def test_forward_returns_patch_tokens_and_fails_on_cpu():
    from m3l_b200._lib import M3LError
    cfg = VD.VTTDinoConfig(depth=1, num_register_tokens=1)
    m = build(cfg, seed=1)
    x = {"image": torch.rand(2, 12, 64, 64), "tactile1": torch.rand(2, 12, 32, 32), "tactile2": torch.rand(2, 12, 32, 32)}
    with pytest.raises(M3LError):
        m(x)
    m = m.to(DEV)
    with torch.no_grad():
        y = m({k: v.to(DEV) for k, v in x.items()})
    assert y.shape == (2, 192, 256)


from tests._golden import VTT_DINO_CASES, VttDinoGolden   # noqa: E402


@pytest.mark.parametrize("name", VTT_DINO_CASES)
def test_forward_features_vs_reference_golden(name):
    """The kernel path against outputs / gradients of the UNMODIFIED reference models/VTT.py::VTT frozen in
    tests/golden/vtt_dino/ (oracle/make_golden_vtt_dino.py)."""
    from m3l_b200.vtt import VTT
    g = VttDinoGolden(name)
    cfg = g.cfg
    m = VTT(image_size=cfg.image_size, tactile_size=cfg.tactile_size, image_patch_size=cfg.image_patch_size,
            tactile_patch_size=cfg.tactile_patch_size, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim,
            num_tactiles=cfg.num_tactiles, image_channels=cfg.image_channels, tactile_channels=cfg.tactile_channels,
            dim_head=cfg.dim_head, num_register_tokens=cfg.num_register_tokens, pos_embed_fn="sinusoidal")
    m.load_state_dict(g.weights(), strict=True)
    m = m.to(DEV)
    masks = g.masks()
    out = m.forward_features({k: v.to(DEV) for k, v in g.inputs().items()}, [mk.to(DEV) for mk in masks] if masks else None)
    for k in ("x_norm_regtokens", "x_norm_patchtokens", "x_prenorm"):
        ref = g.t("out." + k)
        assert out[k].shape == ref.shape
        if ref.numel():
            assert cos(out[k], ref) >= 0.9995, (k, cos(out[k], ref))
    g.objective(out).backward()
    named = dict(m.named_parameters())
    for k, has in g.grad_present().items():
        got = named[k].grad
        assert (got is not None and float(got.abs().max()) > 0) == has, k
    for k, n in g.grad_norms().items():
        if n > 1e-6:
            assert abs(float(named[k].grad.double().norm()) - n) <= 3e-2 * n, (k, float(named[k].grad.double().norm()), n)
    for k, gr in g.full_grads().items():
        assert cos(named[k].grad, gr) >= 0.999, (k, cos(named[k].grad, gr))


def test_vtdino_step_vs_oracle():
    """models/vtdino.py:332-397 on the kernel path (m3l_b200.vtdino.VTDINO): DINO loss, student gradients, the centre
    update and the momentum (EMA) teacher update against the oracle (oracle/vtdino_oracle.py; heads / loss / EMA
    pinned bit-for-bit against the reference files, the backbone against models/VTT.py)."""
    from functools import partial
    from oracle import vtdino_oracle as DO
    from m3l_b200.vtdino import VTDINO, DINOHead
    cfg = VD.VTTDinoConfig(depth=2, num_register_tokens=1, heads=4)
    enc = build(cfg, seed=41)
    torch.manual_seed(5)
    model = VTDINO(enc, partial(DINOHead, out_dim=256, hidden_dim=128, bottleneck_dim=64), num_global_masks=2, num_local_masks=3,
                   moving_average_decay=0.9, teacher_temp=0.05)
    with torch.no_grad():     # make the teacher differ from the student and the centre non-trivial
        for p in model.teacher_encoder.parameters():
            p.add_(torch.randn_like(p) * 0.01)
        model.dino_loss.center.normal_(0, 0.1)
    sds = {k: v.detach().clone() for k, v in model.student_encoder["backbone"].state_dict().items()}
    sdt = {k: v.detach().clone() for k, v in model.teacher_encoder["backbone"].state_dict().items()}
    hs = {k: v.detach().clone() for k, v in model.student_encoder["dino_head"].state_dict().items()}
    ht = {k: v.detach().clone() for k, v in model.teacher_encoder["dino_head"].state_dict().items()}
    center = model.dino_loss.center.detach().clone()
    model = model.to(DEV)
    g = torch.Generator().manual_seed(77)
    B = 4
    x = {"image": torch.rand(B, 12, 64, 64, generator=g), "tactile1": torch.rand(B, 12, 32, 32, generator=g),
         "tactile2": torch.rand(B, 12, 32, 32, generator=g)}
    gm = [torch.stack([torch.randperm(64, generator=g)[:30] for _ in range(B)]) for _ in range(2)]
    lm = [torch.stack([torch.randperm(64, generator=g)[:12] for _ in range(B)]) for _ in range(3)]
    loss = model({k: v.to(DEV) for k, v in x.items()}, [m.to(DEV) for m in gm], [m.to(DEV) for m in lm])
    loss.backward()
    for d in (sds, hs):
        for k in d:
            if d[k].dtype.is_floating_point and k != "pos_embed.frequency_bands":
                d[k].requires_grad_(True)
    lref, t_cls = DO.vtdino_forward(sds, hs, sdt, ht, cfg, x, gm, lm, center, 0.05)
    lref.backward()
    assert abs(float(loss) - float(lref)) <= 1e-2 * abs(float(lref)), (float(loss), float(lref))
    checked = 0
    for prefix, mod, d in (("backbone", model.student_encoder["backbone"], sds), ("head", model.student_encoder["dino_head"], hs)):
        for k, p in mod.named_parameters():
            gr = d[k].grad
            if gr is None or float(gr.abs().max()) == 0.0:
                continue
            assert p.grad is not None and cos(p.grad, gr) >= 0.995, (prefix, k, cos(p.grad, gr) if p.grad is not None else None)
            checked += 1
    assert checked > 30
    for p in model.teacher_encoder.parameters():
        assert p.grad is None
    # centre update (applied lazily at the next softmax_center_teacher) and the EMA teacher update
    model.dino_loss.apply_center_update()
    want_c = DO.center_update(center, t_cls)
    assert cos(model.dino_loss.center, want_c) >= 0.9999
    model.on_train_batch_end()
    for (k, pt), ps in zip(model.teacher_encoder["backbone"].named_parameters(), model.student_encoder["backbone"].parameters()):
        want = DO.ema(sdt[k], ps.detach().cpu(), 0.9)
        assert torch.equal(pt.detach().cpu(), want), k                  # fp32, same operation order: bit-exact
    for (k, pt), ps in zip(model.teacher_encoder["dino_head"].named_parameters(), model.student_encoder["dino_head"].parameters()):
        assert torch.equal(pt.detach().cpu(), DO.ema(ht[k], ps.detach().cpu(), 0.9)), k
    # the updated teacher is what the next forward uses (bf16 shadows refreshed), and training_step runs end to end
    out = model.training_step({k: v.to(DEV) for k, v in x.items()})
    assert out["loss"].requires_grad and out["ssl_loss"] == out["ssl_loss"]
