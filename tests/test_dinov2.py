"""Frozen DINOv2 ViT-S/14-reg image branch + feature concatenation of the DINO-tac-MAE variant
(/root/reference/train_dino_tac_mae.py:29, models/pretrain_models_dino_cat_mae.py:866-904).

CPU: oracle/dinov2_oracle.py against the `transformers` implementation of the same network with random weights (the
parity source SURVEY.md section 8(d) names; the torch.hub weights / code cannot be fetched offline), with and without
position-embedding interpolation; the product module's state_dict matches the torch.hub naming.
GPU: the kernel path (m3l_b200.DinoV2, DinoCatMAEExtractor) against the oracle on the same weights and observations."""
import pytest
import torch

from oracle import dinov2_oracle as DO
from oracle import vtmae_oracle as O


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("image_size", [70, 518])
def test_oracle_matches_transformers_dinov2_with_registers(image_size):
    tr = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = tr.Dinov2WithRegistersConfig(hidden_size=384, num_hidden_layers=2, num_attention_heads=6, mlp_ratio=4, patch_size=14,
                                       image_size=image_size, num_register_tokens=4, layerscale_value=1.0)
    m = tr.Dinov2WithRegistersModel(cfg).eval()
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "lambda1" in n_:
                p.copy_(0.5 + torch.rand_like(p))
            elif n_.endswith("bias") or "token" in n_ or "position" in n_:
                p.copy_(torch.randn_like(p) * 0.1)
    x = torch.rand(3, 3, 70, 70)
    with torch.no_grad():
        out = m(pixel_values=x)
    sd = DO.hf_to_hub_state_dict(m.state_dict())
    assert (DO.dinov2_forward(sd, x) - out.pooler_output).abs().max() < 1e-5
    assert (DO.dinov2_forward(sd, x, return_tokens=True) - out.last_hidden_state).abs().max() < 1e-5
    assert out.last_hidden_state.shape == (3, 30, 384)           # 1 class + 4 register + 25 patch tokens at 70 x 70


def test_product_state_dict_has_the_hub_names():
    from m3l_b200.dinov2 import DinoV2
    m = DinoV2()
    sd = DO.random_state_dict(grid=37)
    sd["mask_token"] = torch.zeros(1, 384)
    assert set(m.state_dict()) == set(sd)
    m.load_state_dict(sd, strict=True)
    assert m.pos_embed.shape == (1, 1370, 384) and len(m.blocks) == 12


# ------------------------------------------------------------------------------------------------ GPU
DEV = "cuda"


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [5, 37])
def test_dinov2_kernel_path_vs_oracle(grid):
    from m3l_b200.dinov2 import DinoV2, mid_frame_view
    from m3l_b200.data import vt_load_lazy
    sd = DO.random_state_dict(grid=grid, seed=1)
    sd["mask_token"] = torch.zeros(1, 384)
    m = DinoV2(img_size=14 * grid)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    gen = torch.Generator().manual_seed(3)
    B, F = 6, 4
    obs = torch.rand(B, F, 70, 70, 3, generator=gen)
    image = obs.permute(0, 1, 4, 2, 3).reshape(B, 3 * F, 70, 70).contiguous()          # the vt_load'ed stacked image
    x = DO.mid_frame(image, F).contiguous()
    ref = DO.dinov2_forward(sd, x)
    ref_tok = DO.dinov2_forward(sd, x, return_tokens=True)
    got = m(x.to(DEV))
    assert got.shape == (B, 384) and got.dtype == torch.float32
    assert _cos(got.cpu(), ref) >= 0.999, _cos(got.cpu(), ref)
    assert (got.cpu() - ref).abs().max() <= 0.06 * ref.abs().max()
    assert _cos(m(x.to(DEV), return_tokens=True).cpu(), ref_tok) >= 0.999
    # the same three channels as a strided slice of the stacked map and as a view into the raw 5-D observation
    got_slice = m(image.to(DEV)[:, 3:6])
    views = vt_load_lazy({"image": obs.to(DEV)}, frame_stack=F)
    got_raw = m(mid_frame_view(views["image"], F))
    assert torch.equal(got_slice, got) and torch.equal(got_raw, got)


@pytest.mark.gpu
def test_dino_cat_mae_extractor_vs_oracle():
    """models/pretrain_models_dino_cat_mae.py:866-904 on the kernel path: (MAE latents | DINO class token) -> mlp."""
    from m3l_b200 import VTT, VTMAE
    from m3l_b200.dinov2 import DinoV2, DinoCatMAEExtractor
    F, dim = 4, 384
    cfg = O.VTMAEConfig(image_size=(70, 70), tactile_size=(70, 70), image_patch_size=14, tactile_patch_size=14, dim=dim, depth=2,
                        heads=4, mlp_dim=2 * dim, decoder_dim=dim, decoder_depth=1, decoder_heads=4, masking_ratio=0.8)
    sd = O.init_state_dict(cfg, seed=5)
    from tests._build import build_product
    mae = build_product(cfg, weights=sd)
    dsd = DO.random_state_dict(grid=5, depth=3, seed=2)
    dsd["mask_token"] = torch.zeros(1, 384)
    dino = DinoV2(img_size=70, depth=3)
    dino.load_state_dict(dsd)
    dino = dino.to(DEV).eval()
    torch.manual_seed(11)
    ext = DinoCatMAEExtractor(None, dino, mae, dim, False, F).to(DEV).eval()
    gen = torch.Generator().manual_seed(4)
    B = 5
    obs = {"image": torch.rand(B, F, 70, 70, 3, generator=gen), "tactile": torch.rand(B, F, 6, 70, 70, generator=gen) * 2 - 1}
    with torch.no_grad():
        got = ext({k: v.clone() for k, v in obs.items()})
    assert got.shape == (B, dim)
    # oracle composition
    o4 = {"image": obs["image"].permute(0, 2, 3, 1, 4).reshape(B, 70, 70, -1), "tactile": obs["tactile"].reshape(B, -1, 70, 70)}
    vt = O.vt_load(o4, frame_stack=F)
    vit_sd = {k[len("vit_layer."):]: v.detach().cpu() for k, v in ext.state_dict().items() if k.startswith("vit_layer.")}
    lat = O.extractor_forward(sd, cfg, vit_sd, {k: v.clone() for k, v in obs.items()}, vision_only_control=False)
    cls = DO.dinov2_forward(dsd, DO.mid_frame(vt["image"], F).contiguous())
    mlp = ext.mlp.cpu().float()
    with torch.no_grad():
        want = mlp(torch.cat((lat, cls), -1))
    ext.mlp.to(DEV)
    assert _cos(got.cpu(), want) >= 0.999, _cos(got.cpu(), want)
