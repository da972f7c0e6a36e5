"""TEST INFRASTRUCTURE (not product code): CPU restatement of the frozen DINOv2 ViT-S/14-with-registers image
branch of the DINO-tac-MAE variant and of the feature concatenation around it.

What the reference does (paths under /root/reference):
  train_dino_tac_mae.py:29-31                      dino = torch.hub.load('facebookresearch/dinov2', 'dinov2_vits14_reg'), frozen
  models/pretrain_models_dino_cat_mae.py:884-889   obs_viso = vt_torch['image'][:, 3*mid-3 : 3*mid] (mid = frame_stack // 2);
                                                   obs_viso = self.dino_model(obs_viso)          -> CLS features (B, 384)
  models/pretrain_models_dino_cat_mae.py:893-904   MAE latents -> vit_layer.transformer -> mean over tokens -> flatten;
                                                   cat((latents, obs_viso), -1) -> self.mlp -> (B, dim)

The DINOv2 network itself is NOT in the reference tree: it is fetched by torch.hub from an un-pinned branch together with
downloaded weights - unobtainable offline.  PARITY PINNING: this restatement follows the published
architecture (ViT-S/14, 12 pre-norm blocks with LayerScale, 6 heads x 64, MLP 1536, LayerNorm eps 1e-6, 4 register
tokens inserted after the class token, position embedding added to class + patch tokens before the registers) and is
pinned against the `transformers` implementation of the same model (Dinov2WithRegistersModel, random weights, the
parity source SURVEY.md section 8(d) names) by tests/test_dinov2.py; the torch.hub code path itself is unpinned (in
particular its position-embedding interpolation for inputs other than the pre-training size).  State-dict names
follow the torch.hub model (`cls_token`, `pos_embed`, `register_tokens`, `patch_embed.proj.*`, `blocks.{i}.norm1.*`,
`.attn.qkv.*`, `.attn.proj.*`, `.ls1.gamma`, `.norm2.*`, `.mlp.fc1.*`, `.mlp.fc2.*`, `.ls2.gamma`, `norm.*`), which
is what a user of the reference holds.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def hf_to_hub_state_dict(hf_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """transformers Dinov2WithRegistersModel.state_dict() -> torch.hub dinov2 naming (q, k, v stacked into qkv)."""
    out = {
        "cls_token": hf_sd["embeddings.cls_token"], "pos_embed": hf_sd["embeddings.position_embeddings"],
        "register_tokens": hf_sd["embeddings.register_tokens"], "mask_token": hf_sd["embeddings.mask_token"],
        "patch_embed.proj.weight": hf_sd["embeddings.patch_embeddings.projection.weight"],
        "patch_embed.proj.bias": hf_sd["embeddings.patch_embeddings.projection.bias"],
        "norm.weight": hf_sd["layernorm.weight"], "norm.bias": hf_sd["layernorm.bias"],
    }
    i = 0
    while f"encoder.layer.{i}.norm1.weight" in hf_sd:
        p, q = f"encoder.layer.{i}.", f"blocks.{i}."
        for nm in ("norm1", "norm2"):
            out[q + nm + ".weight"], out[q + nm + ".bias"] = hf_sd[p + nm + ".weight"], hf_sd[p + nm + ".bias"]
        a = p + "attention.attention."
        out[q + "attn.qkv.weight"] = torch.cat([hf_sd[a + "query.weight"], hf_sd[a + "key.weight"], hf_sd[a + "value.weight"]], 0)
        out[q + "attn.qkv.bias"] = torch.cat([hf_sd[a + "query.bias"], hf_sd[a + "key.bias"], hf_sd[a + "value.bias"]], 0)
        out[q + "attn.proj.weight"], out[q + "attn.proj.bias"] = hf_sd[p + "attention.output.dense.weight"], hf_sd[p + "attention.output.dense.bias"]
        out[q + "ls1.gamma"], out[q + "ls2.gamma"] = hf_sd[p + "layer_scale1.lambda1"], hf_sd[p + "layer_scale2.lambda1"]
        for nm in ("fc1", "fc2"):
            out[q + "mlp." + nm + ".weight"], out[q + "mlp." + nm + ".bias"] = hf_sd[p + "mlp." + nm + ".weight"], hf_sd[p + "mlp." + nm + ".bias"]
        i += 1
    return {k: v.detach().clone() for k, v in out.items()}


def interpolate_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """(1, 1 + g*g, D) table -> (1, 1 + gh*gw, D): class row kept, patch rows resampled bicubically with antialiasing
    in fp32 (as transformers does; identity when the grid already matches)."""
    n_pos = pos_embed.shape[1] - 1
    if n_pos == gh * gw and gh == gw:
        return pos_embed
    g = int(math.sqrt(n_pos))
    D = pos_embed.shape[-1]
    patch = pos_embed[:, 1:].reshape(1, g, g, D).permute(0, 3, 1, 2).to(torch.float32)
    patch = F.interpolate(patch, size=(gh, gw), mode="bicubic", align_corners=False, antialias=True)
    return torch.cat([pos_embed[:, :1], patch.permute(0, 2, 3, 1).reshape(1, gh * gw, D).to(pos_embed.dtype)], 1)


def dinov2_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, patch: int = 14, heads: int = 6, eps: float = 1e-6,
                   return_tokens: bool = False) -> torch.Tensor:
    """x: (B, 3, H, W) fp32 -> normalised class token (B, D)  [return_tokens: all normalised tokens (B, 1+R+N, D)]."""
    B, _, H, W = x.shape
    D = sd["cls_token"].shape[-1]
    tok = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    tok = torch.cat([sd["cls_token"].expand(B, -1, -1), tok], 1) + interpolate_pos_embed(sd["pos_embed"], H // patch, W // patch)
    tok = torch.cat([tok[:, :1], sd["register_tokens"].expand(B, -1, -1), tok[:, 1:]], 1)
    n = tok.shape[1]
    dh = D // heads
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        p = f"blocks.{i}."
        y = F.layer_norm(tok, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, n, 3, heads, dh).permute(2, 0, 3, 1, 4)
        a = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * dh ** -0.5, -1) @ qkv[2]
        a = F.linear(a.transpose(1, 2).reshape(B, n, D), sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        tok = tok + sd[p + "ls1.gamma"] * a
        y = F.layer_norm(tok, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        tok = tok + sd[p + "ls2.gamma"] * y
        i += 1
    tok = F.layer_norm(tok, (D,), sd["norm.weight"], sd["norm.bias"], eps)
    return tok if return_tokens else tok[:, 0]


def mid_frame(image: torch.Tensor, frame_stack: int) -> torch.Tensor:
    """pretrain_models_dino_cat_mae.py:884-889: channels [3*mid-3, 3*mid) of the vt_load'ed image, mid = frame_stack // 2."""
    mid = frame_stack // 2
    return image[:, 3 * mid - 3:3 * mid]


def random_state_dict(dim=384, depth=12, heads=6, patch=14, grid=5, registers=4, mlp_ratio=4, seed=0, ls_init=1.0):
    """Random weights at DINOv2 shapes (LayerScale ~ U[0.5, 1.5] * ls_init so every branch matters in parity tests)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    sd = {"cls_token": r(1, 1, dim, std=0.5), "pos_embed": r(1, 1 + grid * grid, dim, std=0.5), "register_tokens": r(1, registers, dim, std=0.5),
          "patch_embed.proj.weight": r(dim, 3, patch, patch, std=0.05), "patch_embed.proj.bias": r(dim, std=0.1),
          "norm.weight": 1 + r(dim, std=0.1), "norm.bias": r(dim, std=0.1)}
    for i in range(depth):
        p = f"blocks.{i}."
        for nm in ("norm1", "norm2"):
            sd[p + nm + ".weight"], sd[p + nm + ".bias"] = 1 + r(dim, std=0.1), r(dim, std=0.1)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = r(3 * dim, dim, std=0.06), r(3 * dim, std=0.05)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = r(dim, dim, std=0.05), r(dim, std=0.05)
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = r(mlp_ratio * dim, dim, std=0.05), r(mlp_ratio * dim, std=0.05)
        sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = r(dim, mlp_ratio * dim, std=0.03), r(dim, std=0.05)
        sd[p + "ls1.gamma"] = (0.5 + torch.rand(dim, generator=g)) * ls_init
        sd[p + "ls2.gamma"] = (0.5 + torch.rand(dim, generator=g)) * ls_init
    return sd
