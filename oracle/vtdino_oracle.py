"""TEST INFRASTRUCTURE (not product code): CPU restatement of the DINO self-distillation step around the DINO-side
encoder, paths relative to /root/reference:

  DINOHead.forward              tactile_ssl/model/layers/dino_head.py:43-48   (MLP -> L2 normalise -> weight-normed Linear)
  DINOLoss                      tactile_ssl/loss/dino_loss.py:28-101          (softmax_center_teacher, forward, centre EMA)
  update_moving_average         tactile_ssl/utils/ema.py:6-19
  VTDINO.forward                models/vtdino.py:332-397

PINNING: dino_head.py, dino_loss.py and ema.py only need torch, so tests/test_oracle_vs_reference.py loads the
UNMODIFIED files by path and checks these functions bit-for-bit; VTDINO.forward itself is restated on top of
oracle/vtt_dino_oracle.forward_features (pinned separately) - models/vtdino.py needs lightning / hydra / wandb and is
not importable here, so the composition (token reshapes, the list arguments of the loss) is pinned by reading only.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import vtt_dino_oracle as VD


def dino_head_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """DINOHead with nlayers >= 2, no batch-norm: mlp.{0,2,4,...} Linears with GELU between, then normalise and the
    weight-normed last layer (weight = g * v / ||v||_row)."""
    i = 0
    keys = sorted({int(k[len(prefix) + 4:].split(".")[0]) for k in sd if k.startswith(prefix + "mlp.") and k.endswith(".weight")})
    for j, li in enumerate(keys):
        x = F.linear(x, sd[f"{prefix}mlp.{li}.weight"], sd.get(f"{prefix}mlp.{li}.bias"))
        if j + 1 < len(keys):
            x = F.gelu(x)
    x = F.normalize(x, dim=-1, p=2, eps=1e-12)
    g, v = sd[prefix + "last_layer.weight_g"], sd[prefix + "last_layer.weight_v"]
    return F.linear(x, torch._weight_norm(v, g, 0))       # the primitive torch.nn.utils.weight_norm applies


def softmax_center_teacher(t: torch.Tensor, center: torch.Tensor, temp: float) -> torch.Tensor:
    return F.softmax((t - center) / temp, dim=-1)


def dino_loss(student_list: List[torch.Tensor], teacher_list: List[torch.Tensor], student_temp: float = 0.1) -> torch.Tensor:
    total = 0
    for s in student_list:
        lsm = F.log_softmax(s / student_temp, dim=-1)
        for t in teacher_list:
            total = total - torch.sum(t * lsm, dim=-1).mean()
    return total


def center_update(center: torch.Tensor, teacher_output: torch.Tensor, momentum: float = 0.9, world: int = 1) -> torch.Tensor:
    t = torch.sum(teacher_output, dim=0, keepdim=True) / (len(teacher_output) * world)
    return center * momentum + t * (1 - momentum)


def ema(old: torch.Tensor, new: torch.Tensor, beta: float) -> torch.Tensor:
    return old * beta + (1.0 - beta) * new


def vtdino_forward(student_sd, student_head_sd, teacher_sd, teacher_head_sd, cfg: VD.VTTDinoConfig, x: dict,
                   global_masks: List[torch.Tensor], local_masks: List[torch.Tensor], center: torch.Tensor,
                   teacher_temp: float, student_temp: float = 0.1):
    """-> (loss, teacher head output) following models/vtdino.py:332-397."""
    pg, pl = len(global_masks), len(local_masks)
    sg = VD.forward_features(student_sd, cfg, x, global_masks)["x_norm_regtokens"]            # ((p b), 1, c)
    B = sg.shape[0] // pg
    sg = sg.reshape(pg, B, -1).permute(1, 0, 2)
    sl = VD.forward_features(student_sd, cfg, x, local_masks)["x_norm_regtokens"].reshape(pl, B, -1).permute(1, 0, 2)
    s_cls = dino_head_forward(student_head_sd, torch.cat([sg, sl], dim=-2)).permute(1, 0, 2).unsqueeze(2)   # p b 1 c
    with torch.no_grad():
        tg = VD.forward_features(teacher_sd, cfg, x, global_masks)["x_norm_regtokens"]
        t_cls = dino_head_forward(teacher_head_sd, tg)
        t_soft = softmax_center_teacher(t_cls, center, teacher_temp).view(pg, -1, *t_cls.shape[1:])
    return dino_loss(list(s_cls), list(t_soft), student_temp), t_cls
