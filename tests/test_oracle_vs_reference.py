"""CPU, build container only: the oracle against the UNMODIFIED reference module imported through
oracle/stubs (skipped where /root/reference does not exist, e.g. on the GPU box)."""
import pytest
import torch

from oracle import reference_adapter as R
from oracle import vtmae_oracle as O

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("ecm,nt,sincos", [(False, 2, True), (False, 0, True), (True, 2, True), (False, 2, False)])
def test_forward_backward_and_embeddings_match_reference(ecm, nt, sincos):
    cfg = O.VTMAEConfig(early_conv_masking=ecm, num_tactiles=nt, use_sincosmod_encodings=sincos,
                        depth=2, decoder_depth=1)
    mae = R.build_reference_model(cfg, seed=3)
    sd_ref = mae.state_dict()
    mine = O.expand_aliases(O.init_state_dict(cfg))
    assert set(mine) == set(sd_ref)
    assert all(mine[k].shape == sd_ref[k].shape for k in mine)
    sd = O.canonical({k: v.clone() for k, v in sd_ref.items()})
    for k, v in O.position_buffers(cfg).items():
        assert torch.equal(v, sd_ref[k]), k
    g = torch.Generator().manual_seed(7)
    B = 3
    x = {"image": torch.rand(B, 12, 64, 64, generator=g)}
    for i in range(nt):
        x[f"tactile{i + 1}"] = torch.rand(B, 12, 32, 32, generator=g)
    noise = O.tie_free_noise(B, cfg.n_img + nt * cfg.n_tac, g, [64] * (1 + nt))
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    loss = O.vtmae_forward(sd, cfg, x, noise)
    loss.backward()
    with R.injected_noise(R.split_noise(noise, cfg, True, nt > 0)):
        lref = mae(x)
    lref.backward()
    assert torch.equal(loss.detach(), lref.detach())
    for k, p in mae.named_parameters():
        if k not in sd:
            continue
        if p.grad is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
        else:
            assert torch.allclose(sd[k].grad, p.grad, rtol=1e-5, atol=1e-7), k
    with torch.no_grad():
        assert torch.equal(O.vtmae_embeddings(sd, cfg, x), mae.get_embeddings(x, eval=False))
        if nt:
            assert torch.equal(O.vtmae_embeddings(sd, cfg, x, use_tactile=False),
                               mae.get_embeddings(x, eval=False, use_tactile=False))


@pytest.mark.parametrize("nt,sincos,ratio,ecm", [(2, True, None, False), (2, False, 0.5, False), (0, True, 0.8, False),
                                                 (2, True, 0.6, True)])
def test_reconstruct_matches_reference(nt, sincos, ratio, ecm):
    """VTMAE.reconstruct (pretrain_models.py:344-586): per-modality mask counts, rec / masked maps, losses."""
    cfg = O.VTMAEConfig(num_tactiles=nt, use_sincosmod_encodings=sincos, depth=2, decoder_depth=1, early_conv_masking=ecm)
    mae = R.build_reference_model(cfg, seed=5)
    sd = O.canonical({k: v.clone() for k, v in mae.state_dict().items()})
    g = torch.Generator().manual_seed(11)
    B = 3
    x = {"image": torch.rand(B, 12, 64, 64, generator=g)}
    for i in range(nt):
        x[f"tactile{i + 1}"] = torch.rand(B, 12, 32, 32, generator=g)
    noise = O.tie_free_noise(B, cfg.n_img + nt * cfg.n_tac, g, [64] * (1 + nt))
    with torch.no_grad():
        mine = O.vtmae_reconstruct(sd, cfg, x, noise, mask_ratio=ratio)
        with R.injected_noise(R.split_noise(noise, cfg, True, nt > 0)):
            ref = mae.reconstruct({k: v.clone() for k, v in x.items()}, mask_ratio=ratio)
    assert list(mine) == list(ref)
    for k in ref:
        assert torch.equal(mine[k], ref[k]), k


def test_vt_load_matches_reference():
    ref = R.load_reference_module()
    import numpy as np
    rng = np.random.default_rng(0)
    obs = {"image": rng.random((2, 64, 64, 12), dtype=np.float32),
           "tactile": (rng.random((2, 24, 32, 32), dtype=np.float32) * 2 - 1)}
    a = O.vt_load({k: v.copy() for k, v in obs.items()}, frame_stack=4)
    b = ref.vt_load({k: v.copy() for k, v in obs.items()}, frame_stack=4)
    assert set(a) == set(b) == {"image", "tactile1", "tactile2"}
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_train_step_matches_reference_optimizer():
    cfg = O.VTMAEConfig(depth=1, decoder_depth=1)
    mae = R.build_reference_model(cfg, seed=5)
    sd = O.canonical({k: v.clone() for k, v in mae.state_dict().items()})
    g = torch.Generator().manual_seed(11)
    x = {"image": torch.rand(2, 12, 64, 64, generator=g), "tactile1": torch.rand(2, 12, 32, 32, generator=g),
         "tactile2": torch.rand(2, 12, 32, 32, generator=g)}
    noise = O.tie_free_noise(2, 192, g, [64, 64, 64])
    opt = torch.optim.AdamW(mae.parameters(), lr=1e-4)
    st = O.AdamWState()
    for _ in range(2):
        opt.zero_grad()
        with R.injected_noise(R.split_noise(noise, cfg)):
            l = mae(x)
        l.backward()
        torch.nn.utils.clip_grad_norm_(mae.parameters(), 0.5)
        opt.step()
        lo, _, _ = O.train_step(sd, cfg, x, noise, st)
        assert torch.allclose(lo, l.detach(), rtol=1e-6)
    ref_sd = mae.state_dict()
    for k in O.param_keys(sd):
        assert torch.allclose(sd[k].detach(), ref_sd[k], rtol=1e-6, atol=1e-8), k


def test_dino_tac_mae_config_matches_reference():
    """BASELINE.json configs[3], MAE side: 70x70 maps, patch 14, dim 384, tactile-only call (x without 'image':
    pretrain_models.py:148-152), mask 0.8 (train_dino_tac_mae.py:76-80,139-164)."""
    cfg = O.VTMAEConfig(image_size=(70, 70), tactile_size=(70, 70), image_patch_size=14, tactile_patch_size=14,
                        dim=384, depth=2, heads=4, mlp_dim=768, decoder_dim=384, decoder_depth=1, decoder_heads=4,
                        masking_ratio=0.8)
    mae = R.build_reference_model(cfg, seed=9)
    sd = O.canonical({k: v.clone() for k, v in mae.state_dict().items()})
    g = torch.Generator().manual_seed(21)
    B = 3
    x = {f"tactile{i + 1}": torch.rand(B, 12, 70, 70, generator=g) for i in range(2)}
    noise = O.tie_free_noise(B, 2 * cfg.n_tac, g, [cfg.n_tac] * 2)
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    loss = O.vtmae_forward(sd, cfg, x, noise)
    loss.backward()
    # tactile-only: the reference still draws an empty (B, 0) image noise first (:229)
    with R.injected_noise([noise[:, :0]] + R.split_noise(noise, cfg, False, True)):
        lref = mae(x)
    lref.backward()
    assert torch.equal(loss.detach(), lref.detach())
    for k, p in mae.named_parameters():
        if k in sd and p.grad is not None:
            assert torch.allclose(sd[k].grad, p.grad, rtol=1e-5, atol=1e-7), k


@pytest.mark.parametrize("regs,with_masks", [(1, False), (1, True), (0, False), (2, True)])
def test_vtt_dino_forward_features_match_reference(regs, with_masks):
    """models/VTT.py::VTT.forward_features: sinusoidal table, three patch embeddings, shared
    keep-index masks, register tokens, final LayerNorm(eps=1e-6)."""
    from oracle import vtt_dino_oracle as VD
    cfg = VD.VTTDinoConfig(depth=2, num_register_tokens=regs)
    ref = R.build_reference_vtt_dino(cfg, seed=13)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    if regs:                                    # std 1e-6 at init (models/VTT.py:225): make them matter
        with torch.no_grad():
            ref.register_tokens.normal_(0, 0.5)
        sd["register_tokens"] = ref.register_tokens.detach().clone()
    g = torch.Generator().manual_seed(3)
    B = 3
    x = {"image": torch.rand(B, 12, 64, 64, generator=g), "tactile1": torch.rand(B, 12, 32, 32, generator=g),
         "tactile2": torch.rand(B, 12, 32, 32, generator=g)}
    masks = None
    if with_masks:
        masks = [torch.stack([torch.randperm(64, generator=g)[:k] for _ in range(B)]) for k in (40, 40)]
    assert torch.equal(VD.sinusoidal_table(cfg.pos_grid, cfg.dim), ref.pos_embed(torch.device("cpu")))
    for k in sd:
        if sd[k].dtype.is_floating_point and k != "pos_embed.frequency_bands":
            sd[k].requires_grad_(True)
    out = VD.forward_features(sd, cfg, x, masks)
    want = ref.forward_features(x, masks)
    for k in ("x_norm_regtokens", "x_norm_patchtokens", "x_prenorm"):
        assert out[k].shape == want[k].shape and torch.equal(out[k].detach(), want[k].detach()), k
    w = torch.randn(out["x_prenorm"].shape, generator=g)
    (VD.forward_features(sd, cfg, x, masks)["x_norm_patchtokens"].sum() + (out["x_norm_regtokens"] ** 2).sum()).backward()
    o2 = ref.forward_features(x, masks)
    (o2["x_norm_patchtokens"].sum() + (want["x_norm_regtokens"] ** 2).sum()).backward()
    for k, p in ref.named_parameters():
        if p.grad is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
        else:
            assert torch.allclose(sd[k].grad, p.grad, rtol=1e-5, atol=1e-7), k


def test_vtt_dino_product_module_mirrors_reference_state_dict():
    """m3l_b200.vtt.VTT: same state_dict keys / shapes / buffers as models/VTT.py::VTT, same init statistics."""
    from oracle import vtt_dino_oracle as VD
    from m3l_b200.vtt import VTT
    cfg = VD.VTTDinoConfig(num_register_tokens=1)
    ref = R.build_reference_vtt_dino(cfg, seed=0)
    torch.manual_seed(0)
    mine = VTT(image_size=cfg.image_size, tactile_size=cfg.tactile_size, image_patch_size=cfg.image_patch_size,
               tactile_patch_size=cfg.tactile_patch_size, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim,
               num_tactiles=cfg.num_tactiles, image_channels=cfg.image_channels, tactile_channels=cfg.tactile_channels,
               dim_head=cfg.dim_head, num_register_tokens=1, pos_embed_fn="sinusoidal")
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
    assert torch.equal(a["pos_embed.frequency_bands"], b["pos_embed.frequency_bands"])
    assert torch.equal(ref.pos_embed(torch.device("cpu")), mine.pos_embed(torch.device("cpu")))
    w = b["transformer.layers.0.1.net.1.weight"]
    assert abs(float(w.std()) - 0.02) < 2e-3 and float(b["transformer.layers.0.1.net.1.bias"].abs().max()) == 0.0
    mine.load_state_dict(a)


def _load_ref_file(rel, name):
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location(name, str(Path("/root/reference") / rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_dino_head_loss_and_ema_match_reference_files():
    """oracle/vtdino_oracle.py against the UNMODIFIED tactile_ssl/model/layers/dino_head.py, tactile_ssl/loss/dino_loss.py
    and tactile_ssl/utils/ema.py (loaded by path: they only need torch)."""
    from oracle import vtdino_oracle as DO
    head_mod = _load_ref_file("tactile_ssl/model/layers/dino_head.py", "ref_dino_head")
    loss_mod = _load_ref_file("tactile_ssl/loss/dino_loss.py", "ref_dino_loss")
    ema_mod = _load_ref_file("tactile_ssl/utils/ema.py", "ref_ema")
    torch.manual_seed(0)
    head = head_mod.DINOHead(in_dim=64, out_dim=96, hidden_dim=128, bottleneck_dim=32)
    with torch.no_grad():
        head.last_layer.weight_g.mul_(1 + 0.1 * torch.randn_like(head.last_layer.weight_g))
    x = torch.randn(5, 3, 64)
    assert torch.equal(DO.dino_head_forward(head.state_dict(), x), head(x))
    ref_loss = loss_mod.DINOLoss(out_dim=96)
    t_out = torch.randn(10, 1, 96)
    s_list = [torch.randn(5, 1, 96, requires_grad=True) for _ in range(3)]
    for it in range(2):
        t_soft_ref = ref_loss.softmax_center_teacher(t_out, teacher_temp=0.05)
        center_before = ref_loss.center.clone()
        assert torch.equal(DO.softmax_center_teacher(t_out, center_before, 0.05), t_soft_ref)
        ref_loss.update_center(t_out)
        ref_loss.apply_center_update()
        assert torch.equal(DO.center_update(center_before, t_out), ref_loss.center)
        tl = list(t_soft_ref.view(2, -1, 1, 96))
        assert torch.equal(DO.dino_loss(s_list, tl), ref_loss(s_list, tl))
    a, b = torch.nn.Linear(7, 5), torch.nn.Linear(7, 5)
    want = [DO.ema(pa.data.clone(), pb.data.clone(), 0.97) for pb, pa in zip(b.parameters(), a.parameters())]
    ema_mod.update_moving_average(a, b, 0.97)
    for w, pa in zip(want, a.parameters()):
        assert torch.equal(w, pa.data)


def test_sb3_checkpoint_round_trip_with_the_reference_extractor(tmp_path):
    """SB3's CheckpointCallback (utils/callbacks.py:126-133 -> BaseAlgorithm.save) writes a zip whose `policy.pth`
    member is torch.save(policy.state_dict()); the MAE part of a policy lives under `features_extractor.`.  A
    checkpoint written from the UNMODIFIED reference MAEExtractor loads strictly into the product extractor, and one
    written from the product loads strictly back into the reference: saved policies stay interchangeable."""
    import io
    import zipfile
    from m3l_b200 import VTT, VTMAE, MAEExtractor
    ref = R.load_reference_module()
    cfg = O.VTMAEConfig(depth=2, decoder_depth=1)
    ref_mae = R.build_reference_model(cfg, seed=1)
    torch.manual_seed(2)
    ref_ext = ref.MAEExtractor(None, ref_mae, cfg.dim, False, cfg.frame_stack)

    class Policy(torch.nn.Module):             # the slice of an SB3 ActorCriticPolicy that matters here
        def __init__(self, ext):
            super().__init__()
            self.features_extractor = ext
            self.action_net = torch.nn.Linear(cfg.dim, 4)

    def save_sb3_zip(policy, path):
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("data", "{}")
            for member, obj in (("policy.pth", policy.state_dict()), ("policy.optimizer.pth", {}), ("pytorch_variables.pth", {})):
                buf = io.BytesIO()
                torch.save(obj, buf)
                z.writestr(member, buf.getvalue())
            z.writestr("_stable_baselines3_version", "2.1.0")

    def load_policy_pth(path):
        with zipfile.ZipFile(path) as z:
            return torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu")

    enc = VTT(image_size=cfg.image_size, tactile_size=cfg.tactile_size, image_patch_size=cfg.image_patch_size,
              tactile_patch_size=cfg.tactile_patch_size, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim,
              num_tactiles=cfg.num_tactiles, image_channels=cfg.image_channels, tactile_channels=cfg.tactile_channels,
              frame_stack=cfg.frame_stack)
    mae = VTMAE(encoder=enc, decoder_dim=cfg.decoder_dim, masking_ratio=cfg.masking_ratio, decoder_depth=cfg.decoder_depth,
                decoder_heads=cfg.decoder_heads, num_tactiles=cfg.num_tactiles, frame_stack=cfg.frame_stack)
    mine = Policy(MAEExtractor(None, mae, cfg.dim, False, cfg.frame_stack))
    theirs = Policy(ref_ext)
    # reference -> zip -> product
    save_sb3_zip(theirs, tmp_path / "ref.zip")
    sd = load_policy_pth(tmp_path / "ref.zip")
    assert set(sd) == set(mine.state_dict())
    mine.load_state_dict(sd, strict=True)
    for k, v in theirs.state_dict().items():
        assert torch.equal(mine.state_dict()[k], v), k
    # product -> zip -> reference
    with torch.no_grad():
        for p in mine.parameters():
            p.mul_(1.01)
    save_sb3_zip(mine, tmp_path / "mine.zip")
    theirs.load_state_dict(load_policy_pth(tmp_path / "mine.zip"), strict=True)
    for k, v in mine.state_dict().items():
        assert torch.equal(theirs.state_dict()[k], v), k
