"""GPU: vt_load fused into the patch-gather kernels (m3l_patch_source.layout 1, m3l_b200.data.RawMap).

The raw observation tensors (image [B, F, H, W, 3] or [B, H, W, 3F], fp32 or uint8 frames; tactile
[B, F, 3*sensors, h, w] or [B, 3F*sensors, h, w]) are read in place: results must be BIT-IDENTICAL to running the
oracle's vt_load (utils/pretrain_utils.py:7-57, pinned against the reference in test_oracle_vs_reference.py) first and
feeding the kernels fp32 NCHW maps — the arithmetic downstream of the gather is the same code."""
import pytest
import torch

from oracle import vtmae_oracle as O
from tests._build import build_product

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _obs(B, F, sensors, gen, five_d, u8=False):
    img = torch.rand(B, F, 64, 64, 3, generator=gen)
    if u8:
        img = (img * 255).to(torch.uint8)
    tac = torch.rand(B, F, 3 * sensors, 32, 32, generator=gen) * 2 - 1
    if not five_d:        # the 4-D forms the reference hands to vt_load (pretrain_models.py:823-827)
        img = img.permute(0, 2, 3, 1, 4).reshape(B, 64, 64, 3 * F)
        tac = tac.reshape(B, -1, 32, 32)
    return {"image": img.contiguous(), "tactile": tac.contiguous()}


def _oracle_maps(obs, F):
    o = {k: v.clone() for k, v in obs.items()}
    if o["image"].dtype == torch.uint8:
        o["image"] = o["image"].to(torch.float32) / 255          # what a float observation wrapper does before vt_load
    if o["image"].dim() == 5:
        o["image"] = o["image"].permute(0, 2, 3, 1, 4).reshape(o["image"].shape[0], 64, 64, -1)
        o["tactile"] = o["tactile"].reshape(o["tactile"].shape[0], -1, 32, 32)
    return O.vt_load(o, frame_stack=F)


@pytest.mark.parametrize("five_d", [True, False])
@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("F,sensors", [(4, 2), (1, 1), (2, 4)])
def test_vt_load_kernel_bit_exact(five_d, u8, F, sensors):
    from m3l_b200.data import vt_load, vt_load_lazy
    gen = torch.Generator().manual_seed(5)
    obs = _obs(3, F, sensors, gen, five_d, u8)
    want = _oracle_maps(obs, F)
    views = vt_load_lazy({k: v.to(DEV) for k, v in obs.items()}, frame_stack=F)
    assert sorted(views) == sorted(want)
    for k, v in views.items():
        assert v.shape == tuple(want[k].shape)
        assert torch.equal(v.materialize().cpu(), want[k]), k
    if not five_d and not u8:        # the eager drop-in on CUDA observations goes through the same kernel
        got = vt_load({k: v.to(DEV) for k, v in obs.items()}, frame_stack=F)
        assert "tactile" not in got
        for k in want:
            assert torch.equal(got[k].cpu(), want[k]), k


@pytest.mark.parametrize("five_d,u8", [(True, False), (False, False), (True, True)])
def test_patch_kernels_read_raw_observations(five_d, u8):
    from m3l_b200 import ops
    from m3l_b200.data import vt_load_lazy
    gen = torch.Generator().manual_seed(6)
    B, F = 5, 4
    obs = _obs(B, F, 2, gen, five_d, u8)
    maps = {k: v.to(DEV).contiguous() for k, v in _oracle_maps(obs, F).items()}
    views = vt_load_lazy({k: v.to(DEV) for k, v in obs.items()}, frame_stack=F)
    noise = torch.rand(B, 192, generator=gen).to(DEV)
    segs = [(0, 64, 60), (64, 64, 61), (128, 64, 61)]
    masked, unmasked, _ = ops.mask_indices(noise, segs)
    for key, ph, base, ncols_v, col0_v, ncols_m, col0_m, P in (("image", 8, 0, 4, 0, 60, 0, 768), ("tactile", 4, 64, 6, 4, 122, 60, 192)):
        names = ["image"] if key == "image" else ["tactile1", "tactile2"]
        g, b = torch.rand(P, device=DEV) + 0.5, torch.randn(P, device=DEV)
        outs = []
        for src in (maps, views):
            ps = ops.make_patch_source([src[n] for n in names], ph, ph, base)
            a, xhat = ops.patch_layernorm(ps, B, ncols_v, g, b, tok_idx=unmasked, col0=col0_v)
            a_all, _ = ops.patch_layernorm(ps, B, 64 * len(names), g, b, want_xhat=False)
            pred = torch.randn(B * ncols_m, P, generator=torch.Generator().manual_seed(1)).to(DEV)
            loss = torch.zeros(1, device=DEV)
            dcs = torch.zeros(P, device=DEV)
            dp = ops.mse_loss(ps, B, ncols_m, pred, 1.0 / pred.numel(), loss, tok_idx=masked, col0=col0_m, dpred_colsum=dcs)
            outs.append((a, xhat, a_all, dp, loss.clone()))
        for x0, x1 in zip(*outs):
            assert torch.equal(x0, x1), key


@pytest.mark.parametrize("u8", [False, True])
def test_train_step_on_raw_observations_matches_maps(u8):
    """Whole VTMAE step (graph-replayed fused trainer and the autograd path) fed raw 5-D observations through the lazy
    vt_load against the same step fed vt_load'ed maps: same indices, same loss bits, same gradients."""
    from m3l_b200.data import vt_load_lazy
    cfg = O.VTMAEConfig(depth=2, decoder_depth=2)
    sd = O.init_state_dict(cfg, seed=2)
    gen = torch.Generator().manual_seed(9)
    B, F = 6, cfg.frame_stack
    obs = _obs(B, F, 2, gen, True, u8)
    maps = {k: v.to(DEV).contiguous() for k, v in _oracle_maps(obs, F).items()}
    noise = O.tie_free_noise(B, 192, gen, [64] * 3).to(DEV)
    res = []
    for lazy in (False, True):
        mae = build_product(cfg, weights=sd)
        x = vt_load_lazy({k: v.to(DEV) for k, v in obs.items()}, frame_stack=F) if lazy else maps
        loss = mae(x, noise=noise)
        loss.backward()
        grads = {k: p.grad.clone() for k, p in mae.named_parameters() if p.grad is not None}
        mae.initialize_training({"lr": 1e-4, "batch_size": B})
        l2 = [float(mae.train_step(x, noise=noise)) for _ in range(3)]
        res.append((loss.detach().clone(), mae.last_masked_indices.clone(), grads, l2))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for k in res[0][2]:
        assert torch.allclose(res[0][2][k], res[1][2][k], rtol=2e-3, atol=1e-6), k       # fp32 atomics: order only
    for a, b in zip(res[0][3], res[1][3]):
        assert abs(a - b) <= 1e-5 * abs(a)
    # and against the oracle on the same observations
    lref = O.vtmae_forward(sd, cfg, _oracle_maps(obs, F), noise.cpu())
    assert abs(float(res[1][0]) - float(lref)) <= 1e-2 * abs(float(lref))
