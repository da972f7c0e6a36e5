"""TEST INFRASTRUCTURE — CPU oracle for the DINO-side encoder `models/VTT.py::VTT`.
NOT part of the product: only tests/ may import it.

Functional restatement (plain torch fp32 over a flat state_dict) of, paths relative to /root/reference:

  VTT.__init__ / init_weights        models/VTT.py:77-228,801-809
  prepare_tokens_with_masks          models/VTT.py:282-314
  forward_features / forward         models/VTT.py:336-360,424-426
  SinusoidalEmbed                    tactile_ssl/model/layers/patch_embed.py:133-213
  create_ndgrid                      tactile_ssl/utils/__init__.py:39-67
  apply_masks                        tactile_ssl/utils/__init__.py:25-36
  Transformer                        vit-pytorch==1.6.4 (restated in oracle/vtmae_oracle.py; third-party)

PINNING STATUS: everything in /root/reference is pinned bit-for-bit by tests/test_oracle_vs_reference.py
(the unmodified models/VTT.py imported through oracle/stubs); the vit_pytorch Transformer arithmetic is
"parity unpinned" exactly as for the MAE path (package absent, restated from its published algorithm).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from .vtmae_oracle import _pair, layer_norm, linear, patchify, transformer


@dataclass
class VTTDinoConfig:
    image_size: Tuple[int, int] = (64, 64)
    tactile_size: Tuple[int, int] = (32, 32)
    image_patch_size: int = 8
    tactile_patch_size: int = 4
    dim: int = 256
    depth: int = 4
    heads: int = 8
    dim_head: int = 64
    mlp_dim: int = 512
    image_channels: int = 12
    tactile_channels: int = 12
    num_tactiles: int = 2
    num_register_tokens: int = 1

    @property
    def n_img(self):
        (h, w), (p, q) = _pair(self.image_size), _pair(self.image_patch_size)
        return (h // p) * (w // q)

    @property
    def n_tac(self):
        (h, w), (p, q) = _pair(self.tactile_size), _pair(self.tactile_patch_size)
        return (h // p) * (w // q)

    @property
    def pos_grid(self):
        """SinusoidalEmbed([3*H_img, W_img], [p_img, p_img]) (models/VTT.py:195-199): one table for the
        three maps stacked along the rows."""
        (h, w), p = _pair(self.image_size), _pair(self.image_patch_size)[0]
        return (3 * h) // p, w // p


def sinusoidal_table(grid: Tuple[int, int], dim: int) -> torch.Tensor:
    """SinusoidalEmbed.forward (patch_embed.py:188-211) on the integer grid of create_ndgrid(normalized_coords=
    False) (utils/__init__.py:58-66): row-major positions, per axis cat[sin(c*bands), cos(c*bands)], axes
    concatenated, truncated to `dim`."""
    nb = math.ceil(dim / (2 * len(grid)))
    bands = 10000 ** -(torch.linspace(0, 1.0, steps=nb + 1)[:-1])
    freq = torch.stack([bands for _ in grid], dim=0)                       # (ndim, nb)
    axes = [torch.arange(0, r, dtype=torch.float) for r in grid]
    pos = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, len(grid))
    feat = pos[..., None] * freq                                           # (N, ndim, nb)
    enc = torch.cat([torch.sin(feat), torch.cos(feat)], dim=-1).flatten(-2, -1)
    return enc[..., :dim]


EMBED_PREFIXES = ("image_to_patch_embedding", "tactile_to_patch_embedding_1", "tactile_to_patch_embedding_2")


def forward_features(sd: Dict[str, torch.Tensor], cfg: VTTDinoConfig, x: dict, masks: Optional[List[torch.Tensor]] = None):
    """VTT.forward_features (models/VTT.py:336-360).  masks: list of (B, K) int64 keep-index tensors shared by
    the three modalities; the gathered copies are concatenated along the batch, mask-major (apply_masks)."""
    pos = sinusoidal_table(cfg.pos_grid, cfg.dim).float().unsqueeze(0)
    maps = [x["image"], x["tactile1"], x["tactile2"]]
    sizes = [_pair(cfg.image_patch_size), _pair(cfg.tactile_patch_size), _pair(cfg.tactile_patch_size)]
    emb = []
    for m, (p1, p2), pre in zip(maps, sizes, EMBED_PREFIXES):
        t = patchify(m, p1, p2)
        t = layer_norm(t, sd, pre + ".1")
        t = linear(t, sd, pre + ".2")
        emb.append(layer_norm(t, sd, pre + ".3"))
    n1, n2 = emb[0].shape[-2], emb[1].shape[-2]
    emb[0] = emb[0] + pos[:, :n1]                       # the reference's own slice arithmetic (:290-292)
    emb[1] = emb[1] + pos[:, n1:n2 * 2]
    emb[2] = emb[2] + pos[:, n1 * 2:]
    if masks is not None:
        def gather(e):
            return torch.cat([torch.gather(e, -2, mk[..., None].expand(-1, -1, e.shape[-1])) for mk in masks], dim=0)
        emb = [gather(e) for e in emb]
    t = torch.cat(emb, dim=-2)
    R = cfg.num_register_tokens
    if R:
        t = torch.cat((sd["register_tokens"].expand(t.shape[0], -1, -1), t), dim=1)
    t = transformer(t, sd, "transformer", cfg.depth, cfg.heads, cfg.dim_head)
    xn = layer_norm(t, sd, "norm", eps=1e-6)
    return {"x_norm_regtokens": xn[:, :R], "x_norm_patchtokens": xn[:, R:], "x_prenorm": t, "masks": masks}
