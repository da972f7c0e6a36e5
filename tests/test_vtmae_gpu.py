"""GPU parity of the product VTMAE (CUDA kernels through the C-ABI) against
(i) the golden vectors frozen from the unmodified reference and (ii) the CPU oracle run live.

Tolerances (north_star): mask / shuffle indices bit-exact; loss <= 1e-2 relative; gradient cosine
>= 0.999 (bf16 tensor-core compute, fp32 accumulation / LayerNorm / softmax / loss)."""
import pytest
import torch

from oracle import vtmae_oracle as O
from tests._build import build_product
from tests._golden import CASES, Golden

pytestmark = pytest.mark.gpu
DEV = "cuda"

KERNEL_CASES = list(CASES)     # incl. tiny_ecm: early_conv_masking=True (conv stems, loss over all patches)


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


def to_dev(x):
    return {k: v.to(DEV) for k, v in x.items()}


@pytest.mark.parametrize("name", KERNEL_CASES)
def test_forward_backward_vs_reference_golden(name):
    g = Golden(name)
    mae = build_product(g.cfg, weights=g.weights())
    mae.train()
    loss = mae(to_dev(g.inputs()), noise=g.noise().to(DEV))
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.grad_fn is not None
    # integer results: bit-exact
    assert torch.equal(mae.last_masked_indices.cpu(), g.t("masked_indices"))
    assert torch.equal(mae.last_unmasked_indices.cpu(), g.t("unmasked_indices"))
    ref = float(g.t("loss"))
    assert abs(loss.item() - ref) <= 1e-2 * abs(ref), (loss.item(), ref)
    loss.backward()
    named = dict(mae.named_parameters())
    present = g.grad_present()
    for k, has in present.items():
        assert (named[k].grad is not None) == has, f"{k}: grad presence differs from the reference"
    for k, gn in g.grad_norms().items():
        if present[k] and gn > 1e-6:
            mine = float(named[k].grad.double().norm())
            assert abs(mine - gn) <= 3e-2 * gn, (k, mine, gn)
    for k, gr in g.full_grads().items():
        assert cos(named[k].grad, gr) >= 0.999, (k, cos(named[k].grad, gr))


@pytest.mark.parametrize("name", KERNEL_CASES)
def test_embeddings_vs_reference_golden(name):
    g = Golden(name)
    mae = build_product(g.cfg, weights=g.weights())
    with torch.no_grad():
        emb = mae.get_embeddings(to_dev(g.inputs()), eval=False)
    assert emb.shape[1:] == g.t("embeddings").shape[1:] and emb.dtype == torch.float32
    assert mae.training
    ref = g.t("embeddings")
    assert cos(emb[:1], ref) >= 0.9995
    assert (emb[:1].cpu() - ref).abs().max() <= 3e-2 * ref.abs().max()
    if g.has("embeddings_vision_only"):
        with torch.no_grad():
            e2 = mae.get_embeddings(to_dev(g.inputs()), eval=True, use_tactile=False)
        assert not mae.training
        assert cos(e2[:1], g.t("embeddings_vision_only")) >= 0.9995


def _oracle_grads(cfg, sd, x, noise):
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    loss = O.vtmae_forward(sd, cfg, x, noise)
    loss.backward()
    return loss.detach(), {k: sd[k].grad for k in O.param_keys(sd)}


@pytest.mark.parametrize("nt,ecm", [(2, False), (0, False), (2, True), (0, True)])
def test_all_gradients_vs_oracle_canonical(nt, ecm):
    """Every parameter gradient of the canonical model against the CPU oracle, same inputs
    (ecm=True is train.py's default: EarlyCNN conv stems, loss over all patches)."""
    cfg = O.VTMAEConfig(num_tactiles=nt, early_conv_masking=ecm)
    sd = O.init_state_dict(cfg, seed=1)
    gen = torch.Generator().manual_seed(99)
    B = 16
    x = {"image": torch.rand(B, 12, 64, 64, generator=gen)}
    for i in range(nt):
        x[f"tactile{i + 1}"] = torch.rand(B, 12, 32, 32, generator=gen)
    noise = O.tie_free_noise(B, cfg.n_img + nt * cfg.n_tac, gen, [64] * (1 + nt))
    mae = build_product(cfg, weights=sd)
    loss = mae(to_dev(x), noise=noise.to(DEV))
    loss.backward()
    lref, gref = _oracle_grads(cfg, sd, x, noise)
    assert abs(loss.item() - lref.item()) <= 1e-2 * abs(lref.item())
    named = dict(mae.named_parameters(remove_duplicate=False))
    flat_a, flat_b = [], []
    for k, gr in gref.items():
        if gr is None:
            assert named[k].grad is None, k
            continue
        c = cos(named[k].grad, gr)
        assert c >= 0.999, (k, c)
        flat_a.append(named[k].grad.flatten().cpu()); flat_b.append(gr.flatten())
    assert cos(torch.cat(flat_a), torch.cat(flat_b)) >= 0.9995


@pytest.mark.parametrize("B,use_tactile", [(1, True), (7, True), (5, False)])
def test_ragged_batches_vs_oracle(B, use_tactile):
    """Batch sizes that fill no tile (1, 7) and the nt = 2 model called with use_tactile=False
    (MAEExtractor's vision_only_control path, pretrain_models.py:834): loss, indices, gradients."""
    cfg = O.VTMAEConfig(depth=2, decoder_depth=2)
    sd = O.init_state_dict(cfg, seed=8)
    gen = torch.Generator().manual_seed(40 + B)
    x = {"image": torch.rand(B, 12, 64, 64, generator=gen), "tactile1": torch.rand(B, 12, 32, 32, generator=gen),
         "tactile2": torch.rand(B, 12, 32, 32, generator=gen)}
    n = 192 if use_tactile else 64
    noise = O.tie_free_noise(B, n, gen, [64] * (3 if use_tactile else 1))
    mae = build_product(cfg, weights=sd)
    loss = mae(to_dev(x), use_tactile=use_tactile, noise=noise.to(DEV))
    loss.backward()
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    lref = O.vtmae_forward(sd, cfg, x, noise, use_tactile=use_tactile)
    lref.backward()
    assert abs(loss.item() - lref.item()) <= 1e-2 * abs(lref.item())
    named = dict(mae.named_parameters(remove_duplicate=False))
    for k in O.param_keys(sd):
        gr = sd[k].grad
        if gr is None or float(gr.abs().max()) == 0.0:
            assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, k
        else:
            assert cos(named[k].grad, gr) >= 0.999, (k, cos(named[k].grad, gr))


def test_dino_tac_mae_shape_vs_oracle():
    """BASELINE.json configs[3] (DINO-tac-MAE, MAE side): VTT(70x70, patch 14, dim 384, depth 4, heads 4,
    mlp 768, C=12) + VTMAE(r=0.8, decoder_dim 384, depth 3, heads 4) run tactile-only (x without 'image',
    train_dino_tac_mae.py:76-80,139-164): 50 tokens, 40 masked, 10 visible; inner (256) != dim (384)."""
    cfg = O.VTMAEConfig(image_size=(70, 70), tactile_size=(70, 70), image_patch_size=14, tactile_patch_size=14,
                        dim=384, depth=4, heads=4, mlp_dim=768, decoder_dim=384, decoder_depth=3, decoder_heads=4,
                        masking_ratio=0.8)
    sd = O.init_state_dict(cfg, seed=4)
    gen = torch.Generator().manual_seed(17)
    B = 6
    x = {f"tactile{i + 1}": torch.rand(B, 12, 70, 70, generator=gen) for i in range(2)}
    noise = O.tie_free_noise(B, 2 * cfg.n_tac, gen, [cfg.n_tac] * 2)
    mae = build_product(cfg, weights=sd)
    loss = mae(to_dev(x), noise=noise.to(DEV))
    loss.backward()
    assert mae.last_masked_indices.shape == (B, 40) and mae.last_unmasked_indices.shape == (B, 10)
    lref, gref = _oracle_grads(cfg, sd, x, noise)
    assert abs(loss.item() - lref.item()) <= 1e-2 * abs(lref.item())
    named = dict(mae.named_parameters(remove_duplicate=False))
    for k, gr in gref.items():
        if gr is None:
            assert named[k].grad is None, k
        else:
            assert cos(named[k].grad, gr) >= 0.999, (k, cos(named[k].grad, gr))
    with torch.no_grad():
        emb = mae.get_embeddings(to_dev(x), eval=False)
    ref = O.vtmae_embeddings(sd, cfg, x)
    assert emb.shape == ref.shape == (B, 50, 384) and cos(emb, ref) >= 0.9995


def test_embeddings_backward_vs_oracle():
    cfg = O.VTMAEConfig(depth=2)
    sd = O.init_state_dict(cfg, seed=2)
    gen = torch.Generator().manual_seed(5)
    B = 4
    x = {"image": torch.rand(B, 12, 64, 64, generator=gen), "tactile1": torch.rand(B, 12, 32, 32, generator=gen),
         "tactile2": torch.rand(B, 12, 32, 32, generator=gen)}
    w = torch.randn(B, 192, 256, generator=gen)
    mae = build_product(cfg, weights=sd)
    emb = mae.get_embeddings(to_dev(x), eval=False)
    (emb * w.to(DEV)).sum().backward()
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    (O.vtmae_embeddings(sd, cfg, x) * w).sum().backward()
    named = dict(mae.named_parameters(remove_duplicate=False))
    for k in O.param_keys(sd):
        if sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0:
            assert named[k].grad is None, k
        else:
            assert cos(named[k].grad, sd[k].grad) >= 0.999, (k, cos(named[k].grad, sd[k].grad))


def test_standalone_transformer_matches_oracle():
    """MAEExtractor's extra block (pretrain_models.py:807-817,836): Transformer(dim, 1, 4, 64, 2*dim)."""
    from m3l_b200 import Transformer
    torch.manual_seed(3)
    t = Transformer(256, 1, 4, 64, 512).to(DEV)
    sd = {"transformer." + k: v.detach().cpu().clone() for k, v in t.state_dict().items()}
    x = torch.randn(3, 192, 256)
    w = torch.randn(3, 192, 256)
    xg = x.to(DEV).requires_grad_(True)
    y = t(xg)
    (y * w.to(DEV)).sum().backward()
    xr = x.clone().requires_grad_(True)
    for v in sd.values():
        v.requires_grad_(True)
    yr = O.transformer(xr, sd, "transformer", 1, 4, 64)
    (yr * w).sum().backward()
    assert cos(y, yr.detach()) >= 0.9995
    assert cos(xg.grad, xr.grad) >= 0.999
    for k, p in t.named_parameters():
        assert cos(p.grad, sd["transformer." + k].grad) >= 0.999, k


@pytest.mark.parametrize("nt,ratio,ecm", [(2, None, False), (2, 0.5, False), (0, 0.8, False), (2, 0.6, True)])
def test_reconstruct_vs_oracle(nt, ratio, ecm):
    """VTMAE.reconstruct (pretrain_models.py:344-586): same dict, same masked patches, reconstruction close."""
    cfg = O.VTMAEConfig(num_tactiles=nt, depth=2, decoder_depth=2, early_conv_masking=ecm)
    sd = O.init_state_dict(cfg, seed=6)
    gen = torch.Generator().manual_seed(8)
    B = 3
    x = {"image": torch.rand(B, 12, 64, 64, generator=gen)}
    for i in range(nt):
        x[f"tactile{i + 1}"] = torch.rand(B, 12, 32, 32, generator=gen)
    noise = O.tie_free_noise(B, cfg.n_img + nt * cfg.n_tac, gen, [64] * (1 + nt))
    mae = build_product(cfg, weights=sd)
    out = mae.reconstruct(to_dev(x), mask_ratio=ratio, noise=noise.to(DEV))
    with torch.no_grad():
        ref = O.vtmae_reconstruct(sd, cfg, x, noise, mask_ratio=ratio)
    assert list(out) == list(ref)
    for k in ref:
        a, b = out[k].cpu(), ref[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if k.endswith("_masked"):
            assert torch.equal(a, b), k                       # pure data movement + constants: bit-exact
        elif k.startswith("recon_loss"):
            assert abs(float(a) - float(b)) <= 1e-2 * abs(float(b)), (k, float(a), float(b))
        else:
            assert cos(a, b) >= 0.9995 and (a - b).abs().max() < 5e-2, (k, cos(a, b), float((a - b).abs().max()))


def test_autograd_graph_path_matches_eager():
    """loss = mae(x); loss.backward() replayed from CUDA graphs (the default) against the eager kernel sequence:
    identical loss, indices and gradients over several calls with fresh inputs, .grad never aliasing the graph's
    static buffers, and the automatic eager fallback for a forward issued while the previous one is unconsumed."""
    cfg = O.VTMAEConfig(depth=2, decoder_depth=2)
    sd = O.init_state_dict(cfg, seed=3)
    mae_g = build_product(cfg, weights=sd)
    mae_e = build_product(cfg, weights=sd)
    mae_e.use_cuda_graph = False
    gen = torch.Generator().manual_seed(77)
    B = 6
    kept = None
    for it in range(3):
        x = {"image": torch.rand(B, 12, 64, 64, generator=gen), "tactile1": torch.rand(B, 12, 32, 32, generator=gen),
             "tactile2": torch.rand(B, 12, 32, 32, generator=gen)}
        noise = O.tie_free_noise(B, 192, gen, [64] * 3)
        for m in (mae_g, mae_e):
            m.zero_grad(set_to_none=True)
        lg = mae_g(to_dev(x), noise=noise.to(DEV)); lg.backward(torch.tensor(0.5 + it, device=DEV))
        le = mae_e(to_dev(x), noise=noise.to(DEV)); le.backward(torch.tensor(0.5 + it, device=DEV))
        assert torch.equal(lg.detach(), le.detach())
        assert torch.equal(mae_g.last_masked_indices, mae_e.last_masked_indices)
        ge, gg = dict(mae_e.named_parameters()), dict(mae_g.named_parameters())
        for k, p in ge.items():
            assert (p.grad is None) == (gg[k].grad is None), k
            if p.grad is not None:
                assert torch.allclose(gg[k].grad, p.grad, rtol=2e-3, atol=1e-6), k     # fp32 atomics: order only
        if kept is None:
            kept = {k: p.grad.clone() for k, p in gg.items() if p.grad is not None}
            held = {k: p.grad for k, p in gg.items() if p.grad is not None}
            for m in (mae_g,):
                m.zero_grad(set_to_none=True)
        else:
            for k in kept:                        # the tensors autograd handed out earlier were not overwritten
                assert torch.equal(held[k], kept[k]), k
    # two grad-enabled forwards of the same shape before any backward (SAC with a shared extractor, two losses): the
    # second call must not clobber the first one's saved activations -> it takes the eager path; both backwards work
    x1 = to_dev(x)
    mae_g.zero_grad(set_to_none=True)
    l1 = mae_g(x1, noise=noise.to(DEV))
    l2 = mae_g(x1, noise=noise.to(DEV))
    assert torch.equal(l1.detach(), l2.detach())
    l1.backward()
    g1 = {k: p.grad.clone() for k, p in mae_g.named_parameters() if p.grad is not None}
    l2.backward()
    for k, p in mae_g.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, 2 * g1[k], rtol=2e-3, atol=1e-6), k
    # a forward whose result is dropped without backward does not block the graph path for ever
    mae_g(x1, noise=noise.to(DEV))
    l3 = mae_g(x1, noise=noise.to(DEV))
    assert type(l3.grad_fn).__name__.startswith("_MAEGraphFn")
    l3.backward()


def test_mae_extractor_rollout_graph_matches_eager():
    """Rollout-time inference (torch.no_grad) replays a CUDA graph cached per observation shape: same features as the
    eager chain, for new observations and after a parameter update, numpy observations included."""
    from m3l_b200 import MAEExtractor
    cfg = O.VTMAEConfig(depth=2)
    sd = O.init_state_dict(cfg, seed=4)
    gen = torch.Generator().manual_seed(16)
    mae = build_product(cfg, weights=sd)
    ext = MAEExtractor(None, mae, cfg.dim, False, cfg.frame_stack).to(DEV)

    def obs(B):
        return {"image": torch.rand(B, cfg.frame_stack, 64, 64, 3, generator=gen),
                "tactile": torch.rand(B, cfg.frame_stack, 6, 32, 32, generator=gen) * 2 - 1}

    for B in (2, 5, 2):
        o = obs(B)
        with torch.no_grad():
            got = ext({k: v.numpy() for k, v in o.items()})           # SB3 hands numpy arrays over
            ext.use_cuda_graph = False
            want = ext({k: v.to(DEV) for k, v in o.items()})
            ext.use_cuda_graph = True
        assert got.shape == (B, cfg.dim) and torch.equal(got, want)
    with torch.no_grad():
        for p in ext.parameters():
            p.mul_(1.01)                                              # e.g. an optimizer step between rollouts
        o = obs(2)
        got = ext({k: v.to(DEV) for k, v in o.items()})
        ext.use_cuda_graph = False
        want = ext({k: v.to(DEV) for k, v in o.items()})
    assert torch.equal(got, want)
    # training-time call (gradients through the extractor): graph replay against the eager chain
    w = torch.randn(4, cfg.dim, generator=gen).to(DEV)
    grads = []
    for use_graph in (True, False, True):
        ext.use_cuda_graph = use_graph
        ext.zero_grad(set_to_none=True)
        o = obs(4) if use_graph is True and not grads else o
        f = ext({k: v.to(DEV) for k, v in o.items()})
        assert f.requires_grad and mae.training
        (f * w).sum().backward()
        grads.append((f.detach().clone(), {k: p.grad.clone() for k, p in ext.named_parameters() if p.grad is not None}))
    for f, g in grads[1:]:
        assert torch.equal(f, grads[0][0]) and g.keys() == grads[0][1].keys()
        for k in g:
            assert torch.allclose(g[k], grads[0][1][k], rtol=2e-3, atol=1e-6), k


@pytest.mark.parametrize("vision_only", [False, True])
def test_mae_extractor_vs_oracle(vision_only):
    """MAEExtractor.forward (pretrain_models.py:819-841): raw 5-D observations -> (B, dim) features, and the
    gradients the PPO / SAC losses send through it into the extra block and the MAE encoder."""
    from m3l_b200 import MAEExtractor
    cfg = O.VTMAEConfig(depth=2)
    sd = O.init_state_dict(cfg, seed=4)
    gen = torch.Generator().manual_seed(6)
    B, F_ = 3, cfg.frame_stack
    obs = {"image": torch.rand(B, F_, 64, 64, 3, generator=gen),
           "tactile": torch.rand(B, F_, 6, 32, 32, generator=gen) * 2 - 1}
    mae = build_product(cfg, weights=sd)
    torch.manual_seed(9)
    ext = MAEExtractor(None, mae, cfg.dim, vision_only, F_).to(DEV)
    assert ext.features_dim == cfg.dim
    sd_vit = {"transformer." + k: v.detach().cpu().clone() for k, v in ext.vit_layer.transformer.state_dict().items()}
    w = torch.randn(B, cfg.dim, generator=gen)
    feats = ext({k: v.clone().to(DEV) for k, v in obs.items()})
    assert feats.shape == (B, cfg.dim) and feats.dtype == torch.float32 and mae.training
    (feats * w.to(DEV)).sum().backward()
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    for v in sd_vit.values():
        v.requires_grad_(True)
    ref = O.extractor_forward(sd, cfg, sd_vit, {k: v.clone() for k, v in obs.items()}, vision_only_control=vision_only)
    (ref * w).sum().backward()
    assert cos(feats, ref.detach()) >= 0.9995
    for k, p in ext.vit_layer.transformer.named_parameters():
        assert cos(p.grad, sd_vit["transformer." + k].grad) >= 0.999, k
    assert all(p.grad is None for k, p in ext.vit_layer.named_parameters() if not k.startswith("transformer."))
    named = dict(mae.named_parameters(remove_duplicate=False))
    for k in O.param_keys(sd):
        if sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0:
            assert named[k].grad is None, k
        else:
            assert cos(named[k].grad, sd[k].grad) >= 0.999, (k, cos(named[k].grad, sd[k].grad))
    # rollout use: no autograd
    with torch.no_grad():
        f2 = ext({k: v.clone().to(DEV) for k, v in obs.items()})
    assert torch.equal(f2, feats.detach())


@pytest.mark.parametrize("use_graph,ecm", [(False, False), (True, False), (True, True)])
def test_fused_train_steps_vs_oracle(use_graph, ecm):
    """zero_grad + fwd + bwd + clip(0.5) + AdamW (pretrain_models.py:707-711): loss trajectory and
    updated weights against the oracle's restatement of torch AdamW, 3 steps."""
    from m3l_b200.trainer import FusedTrainer
    cfg = O.VTMAEConfig(early_conv_masking=ecm)
    sd = O.init_state_dict(cfg, seed=4)
    gen = torch.Generator().manual_seed(17)
    B = 8
    x = {"image": torch.rand(B, 12, 64, 64, generator=gen), "tactile1": torch.rand(B, 12, 32, 32, generator=gen),
         "tactile2": torch.rand(B, 12, 32, 32, generator=gen)}
    noise = O.tie_free_noise(B, 192, gen, [64, 64, 64])
    mae = build_product(cfg, weights=sd)
    mae.initialize_training({"lr": 1e-4, "batch_size": B})
    mae._trainer.use_graph = use_graph
    st = O.AdamWState()
    xd, nd = to_dev(x), noise.to(DEV)
    w0 = {k: v.detach().clone() for k, v in sd.items()}
    for it in range(3):
        l = mae.train_step(xd, noise=nd)
        lo, norm, _ = O.train_step(sd, cfg, x, noise, st)
        assert abs(l.item() - lo.item()) <= 1e-2 * abs(lo.item()), (it, l.item(), lo.item())
        assert abs(mae._trainer.state[2].item() - norm.item()) <= 3e-2 * norm.item()
    named = dict(mae.named_parameters())
    keys = ["decoder.layers.0.1.net.1.weight", "to_pixels.weight", "mask_token", "encoder.transformer.layers.0.0.to_qkv.weight"]
    if ecm:
        keys += ["early_conv_vision.conv1.weight", "early_conv_tactile.conv3.weight", "early_conv_vision.conv4.bias"]
    for k in keys:
        upd_ref = sd[k].detach() - w0[k]
        upd = named[k].detach().cpu() - w0[k]
        assert cos(upd, upd_ref) >= 0.99, (k, cos(upd, upd_ref))
    # parameters the reference leaves without gradient are untouched (no weight decay either)
    assert torch.equal(named["encoder.pos_embedding"].detach().cpu(), w0["encoder.pos_embedding"])
    assert mae._trainer.state[0].item() == 3.0


def test_cpu_module_fails_loudly():
    from m3l_b200 import M3LError
    mae = build_product(O.VTMAEConfig(depth=1, decoder_depth=1), device="cpu")
    with pytest.raises(M3LError):
        mae({"image": torch.zeros(1, 12, 64, 64)})


@pytest.mark.parametrize("ecm", [False, True])
def test_joint_mae_and_extractor_pass_matches_separate_passes(ecm):
    """SURVEY.md 8(f)-2 / ppo_mae.py:255-283: `mae(x).backward()` and the extractor forward / backward of
    `evaluate_actions` over the same minibatch as ONE pass (shared patch embedding): same loss, same features, and the
    same accumulated gradients as the two separate passes (which are checked against the oracle above)."""
    from m3l_b200 import MAEExtractor
    from m3l_b200.data import vt_load_lazy
    cfg = O.VTMAEConfig(depth=2, decoder_depth=2, early_conv_masking=ecm)
    sd = O.init_state_dict(cfg, seed=8)
    gen = torch.Generator().manual_seed(21)
    B = 6
    mae = build_product(cfg, weights=sd)
    ext = MAEExtractor(None, mae, cfg.dim, False, cfg.frame_stack).to(DEV)
    obs = {"image": torch.rand(B, cfg.frame_stack, 64, 64, 3, generator=gen).to(DEV),
           "tactile": (torch.rand(B, cfg.frame_stack, 6, 32, 32, generator=gen) * 2 - 1).to(DEV)}
    noise = O.tie_free_noise(B, 192, gen, [64] * 3).to(DEV)
    w = torch.randn(B, cfg.dim, generator=gen).to(DEV)

    def grads():
        return {k: p.grad.clone() for k, p in ext.named_parameters() if p.grad is not None}

    # separate passes (the reference's sequence)
    ext.zero_grad(set_to_none=True)
    l_sep = mae(vt_load_lazy(obs, frame_stack=cfg.frame_stack), noise=noise)
    l_sep.backward()
    f_sep = ext(obs)
    (f_sep * w).sum().backward()
    g_sep = grads()
    for rep in range(3):          # graph capture, replay, replay
        ext.zero_grad(set_to_none=True)
        l_j = ext.joint_mae_loss(obs, noise=noise)
        f_j = ext(obs)
        assert f_j.grad_fn is l_j.grad_fn or type(f_j.grad_fn).__name__ != type(f_sep.grad_fn).__name__
        assert torch.equal(l_j.detach(), l_sep.detach())
        assert torch.allclose(f_j.detach(), f_sep.detach(), rtol=0, atol=0)
        ((f_j * w).sum() + l_j).backward()
        g_j = grads()
        assert g_j.keys() == g_sep.keys()
        for k in g_sep:
            c = cos(g_j[k], g_sep[k])
            assert c >= 0.9995, (k, c)
            assert torch.allclose(g_j[k], g_sep[k], rtol=5e-2, atol=2e-3 * float(g_sep[k].abs().max()) + 1e-7), k
    # outstanding joint pass -> the next one takes the eager path; both backwards work
    ext.zero_grad(set_to_none=True)
    l1 = ext.joint_mae_loss(obs, noise=noise)
    f1 = ext(obs)
    l2 = ext.joint_mae_loss(obs, noise=noise)
    f2 = ext(obs)
    assert torch.equal(l1.detach(), l2.detach())
    ((f1 * w).sum() + l1).backward()
    ((f2 * w).sum() + l2).backward()
    g2 = grads()
    for k in g_sep:
        assert cos(g2[k], g_sep[k]) >= 0.9995, k


def test_ppo_mae_update_sequence_with_shared_adam():
    """The MAE-related lines of one PPO_MAE minibatch update (/root/reference/models/ppo_mae.py:236-283, non-separate
    optimizer): policy.optimizer.zero_grad(); mae_loss = mae(vt_load(obs)); mae_loss.backward(); features =
    extractor(obs) -> policy loss; loss.backward(); clip_grad_norm_; policy.optimizer.step() with ONE torch Adam over
    all policy parameters (extractor incl. the MAE).  Product modules driven exactly like that for three updates must
    track the oracle driven the same way: losses within 1e-2, per-tensor parameter updates in the same direction."""
    from m3l_b200 import MAEExtractor
    from m3l_b200.data import vt_load
    cfg = O.VTMAEConfig(depth=2, decoder_depth=2)
    sd = O.init_state_dict(cfg, seed=12)
    gen = torch.Generator().manual_seed(33)
    B, F = 8, cfg.frame_stack
    mae = build_product(cfg, weights=sd)
    ext = MAEExtractor(None, mae, cfg.dim, False, F).to(DEV)
    head = torch.nn.Linear(cfg.dim, 3).to(DEV)                       # stands in for the actor / critic heads
    opt = torch.optim.Adam(list(ext.parameters()) + list(head.parameters()), lr=3e-4)
    # oracle side: same weights as leaf tensors
    vit_sd = {k[len("vit_layer."):]: v.detach().cpu().clone().requires_grad_(True) for k, v in ext.state_dict().items()
              if k.startswith("vit_layer.transformer")}
    osd = {k: v.clone() for k, v in sd.items()}
    for k in O.param_keys(osd):
        osd[k].requires_grad_(True)
    ohead = torch.nn.Linear(cfg.dim, 3)
    ohead.load_state_dict({k: v.cpu() for k, v in head.state_dict().items()})
    oparams = [osd[k] for k in O.param_keys(osd)] + list(vit_sd.values()) + list(ohead.parameters())
    oopt = torch.optim.Adam(oparams, lr=3e-4)
    for it in range(3):
        obs = {"image": torch.rand(B, F, 64, 64, 3, generator=gen), "tactile": torch.rand(B, F, 6, 32, 32, generator=gen) * 2 - 1}
        noise = O.tie_free_noise(B, 192, gen, [64] * 3)
        tgt = torch.randn(B, 3, generator=gen)
        # ---- product, the reference's call sequence
        observations = {k: v.to(DEV) for k, v in obs.items()}
        observations["image"] = observations["image"].permute(0, 2, 3, 1, 4).reshape(B, 64, 64, -1)
        observations["tactile"] = observations["tactile"].reshape(B, -1, 32, 32)
        opt.zero_grad()
        x = vt_load({k: v.clone() for k, v in observations.items()}, frame_stack=F)
        mae_loss = mae(x, noise=noise.to(DEV))
        mae_loss.backward()
        feats = ext(observations)
        loss = ((head(feats) - tgt.to(DEV)) ** 2).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(ext.parameters()) + list(head.parameters()), 0.5)
        opt.step()
        # ---- oracle, same sequence
        oopt.zero_grad()
        o4 = {"image": obs["image"].permute(0, 2, 3, 1, 4).reshape(B, 64, 64, -1), "tactile": obs["tactile"].reshape(B, -1, 32, 32)}
        ox = O.vt_load({k: v.clone() for k, v in o4.items()}, frame_stack=F)
        ol = O.vtmae_forward(osd, cfg, ox, noise)
        ol.backward()
        of = O.extractor_forward(osd, cfg, vit_sd, {k: v.clone() for k, v in obs.items()}, vision_only_control=False)
        oloss = ((ohead(of) - tgt) ** 2).mean()
        oloss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in oparams if p.grad is not None], 0.5)
        oopt.step()
        assert abs(float(mae_loss) - float(ol)) <= 1e-2 * abs(float(ol)), (it, float(mae_loss), float(ol))
        assert abs(float(loss) - float(oloss)) <= 2e-2 * abs(float(oloss)) + 1e-4, (it, float(loss), float(oloss))
    named = dict(mae.named_parameters(remove_duplicate=False))
    # Adam's update is m / (sqrt(v) + eps): in the first steps close to sign(g) per element, so elements whose gradient
    # is at the bf16 noise level may flip - the update DIRECTIONS agree per tensor (>= 0.95) and on average (>= 0.99)
    cs = []
    for k in O.param_keys(osd):
        if k in named and osd[k].grad is not None:
            d_ref = osd[k].detach() - sd[k]
            d_got = named[k].detach().cpu() - sd[k]
            if float(d_ref.norm()) > 0:
                cs.append(cos(d_got, d_ref))
                assert cs[-1] >= 0.95, (k, cs[-1])
    assert len(cs) > 50 and sum(cs) / len(cs) >= 0.99, (len(cs), sum(cs) / len(cs))
