"""DINO self-distillation around the DINO-side encoder (`m3l_b200.vtt.VTT`): student / teacher backbones and heads,
the centred-softmax cross-entropy and the momentum teacher — the compute path of `models/vtdino.py::VTDINO`.

Reference (paths under /root/reference):
  models/vtdino.py:28-127       constructor: student = {backbone, dino_head}, teacher = deep copy (no grad), DINOLoss
  models/vtdino.py:159-173      on_train_batch_end: teacher temperature schedule, EMA teacher update
  models/vtdino.py:212-330      block-mask sampling (host side)
  models/vtdino.py:332-397      forward(x, global_masks, local_masks) -> loss
  tactile_ssl/model/layers/dino_head.py:12-67   DINOHead (MLP -> L2 normalise -> weight-normed Linear without bias)
  tactile_ssl/loss/dino_loss.py:10-101          DINOLoss (softmax_center_teacher, forward, centre EMA with all-reduce)
  tactile_ssl/utils/ema.py:6-19                 update_moving_average

The backbones run on the sm_100a kernel path (vtt.py: one autograd.Function over patchify+LN, tcgen05 GEMMs, fused
attention, ...).  The Linear layers of the heads (the last one is bottleneck_dim x out_dim, tens of thousands of
prototypes) go through the same tcgen05 GEMM (forward, dgrad and wgrad) via `_KernelLinear`; the teacher update is ONE
kernel over the flat parameter arenas (`m3l_ema_update`).  GELU / L2-normalise / (log-)softmax of the heads and the loss
act on (views x batch, out_dim) matrices and stay torch elementwise / reduction ops.  Lightning plumbing (optimizer and
scheduler construction, wandb logging, online probes) is out of scope.
"""
from __future__ import annotations

import copy
import math
from functools import partial
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn
from torch.nn.init import trunc_normal_
from torch.nn.utils import weight_norm

from . import engine, ops
from ._lib import M3LError


class _KernelLinear(torch.autograd.Function):
    """y = x W^T (+ b) on the tcgen05 GEMM: bf16 operands, fp32 accumulation and output; backward = dgrad + wgrad GEMMs."""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1])
        xb, wb = x2.to(torch.bfloat16).contiguous(), w.to(torch.bfloat16).contiguous()
        y = ops.gemm(xb, wb, bias=b.float().contiguous() if b is not None else None, out_dtype=torch.float32)
        ctx.save_for_backward(xb, wb)
        ctx.has_bias, ctx.shape = b is not None, x.shape
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        xb, wb = ctx.saved_tensors
        g = gy.reshape(-1, gy.shape[-1]).to(torch.bfloat16).contiguous()
        dx = ops.gemm(g, wb.t().contiguous(), out_dtype=torch.float32).reshape(ctx.shape) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(wb.shape, dtype=torch.float32, device=wb.device)
            engine.wgrad(g, xb, dw)
        db = gy.reshape(-1, gy.shape[-1]).sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def _klinear(x, w, b):
    if not x.is_cuda:
        raise M3LError("m3l_b200.vtdino: CUDA tensors required (no CPU fallback)")
    if x.shape[-1] % 8 or w.shape[0] % 8:
        return F.linear(x, w, b)          # shapes the GEMM cannot tile (never the case for the DINO head dimensions)
    return _KernelLinear.apply(x, w, b)


class DINOHead(nn.Module):
    """tactile_ssl/model/layers/dino_head.py:12-67 - same constructor, module tree and state_dict
    (mlp.{0,2,4}.*, last_layer.weight_g / weight_v)."""

    def __init__(self, in_dim, out_dim, use_bn=False, nlayers=3, hidden_dim=2048, bottleneck_dim=256, mlp_bias=True):
        super().__init__()
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim, bias=mlp_bias)
        else:
            layers = [nn.Linear(in_dim, hidden_dim, bias=mlp_bias)]
            if use_bn:
                layers.append(nn.BatchNorm1d(hidden_dim))
            layers.append(nn.GELU())
            for _ in range(nlayers - 2):
                layers.append(nn.Linear(hidden_dim, hidden_dim, bias=mlp_bias))
                if use_bn:
                    layers.append(nn.BatchNorm1d(hidden_dim))
                layers.append(nn.GELU())
            layers.append(nn.Linear(hidden_dim, bottleneck_dim, bias=mlp_bias))
            self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)
        self.last_layer = weight_norm(nn.Linear(bottleneck_dim, out_dim, bias=False))
        self.last_layer.weight_g.data.fill_(1)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        mods = [self.mlp] if isinstance(self.mlp, nn.Linear) else list(self.mlp)
        for m in mods:
            x = _klinear(x, m.weight, m.bias) if isinstance(m, nn.Linear) else m(x)
        eps = 1e-6 if x.dtype == torch.float16 else 1e-12
        x = F.normalize(x, dim=-1, p=2, eps=eps)
        g, v = self.last_layer.weight_g, self.last_layer.weight_v        # weight = g * v / ||v|| (per output row)
        w = torch._weight_norm(v, g, 0)
        return _klinear(x, w, None)


class DINOLoss(nn.Module):
    """Centred-softmax cross-entropy of tactile_ssl/loss/dino_loss.py:10-101 (same constructor, `center` buffer and
    method names; the Sinkhorn-Knopp teacher is not used by VTDINO and not provided).  The centre update is lazy, as in
    the reference: `update_center` only starts the (all-reduced) batch sum, the EMA is applied at the next
    `softmax_center_teacher`."""

    def __init__(self, out_dim, student_temp=0.1, center_momentum=0.9):
        super().__init__()
        self.student_temp, self.center_momentum = student_temp, center_momentum
        self.register_buffer("center", torch.zeros(1, out_dim))
        self._pending = None              # (batch sum of teacher outputs, rows summed, async all-reduce handle)

    @staticmethod
    def _world():
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    @torch.no_grad()
    def apply_center_update(self):
        if self._pending is None:
            return
        total, rows, handle = self._pending
        if handle is not None:
            handle.wait()
        batch_center = total / (rows * self._world())
        self.center = self.center * self.center_momentum + batch_center * (1 - self.center_momentum)
        self._pending = None

    @torch.no_grad()
    def softmax_center_teacher(self, teacher_output, teacher_temp):
        self.apply_center_update()
        return F.softmax((teacher_output - self.center) / teacher_temp, dim=-1)

    @torch.no_grad()
    def update_center(self, teacher_output):
        total = torch.sum(teacher_output, dim=0, keepdim=True)
        handle = dist.all_reduce(total, async_op=True) if self._world() > 1 else None
        self._pending = (total, len(teacher_output), handle)

    def forward(self, student_output_list, teacher_out_softmaxed_centered_list):
        """-sum over (student view, teacher view) pairs of mean_b sum_k t_k log softmax(s / student_temp)_k."""
        loss = 0
        for s_out in student_output_list:
            log_p = F.log_softmax(s_out / self.student_temp, dim=-1)
            for t_prob in teacher_out_softmaxed_centered_list:
                loss = loss - torch.sum(t_prob * log_p, dim=-1).mean()
        return loss


def update_moving_average(ma_model: nn.Module, current_model: nn.Module, beta: float) -> None:
    """tactile_ssl/utils/ema.py:13-19.  Backbones whose parameters live in flat arenas (m3l_b200.vtt.VTT) are updated
    by one `m3l_ema_update` kernel per arena; remaining parameters (heads, the position table) pairwise with the same
    kernel."""
    done = set()
    mods_t, mods_s = dict(ma_model.named_modules()), dict(current_model.named_modules())
    for name, mt in mods_t.items():
        ms = mods_s.get(name)
        at, as_ = getattr(mt, "_arena", None), getattr(ms, "_arena", None)
        if at is not None and as_ is not None and at.names == as_.names and at.is_bound() and as_.is_bound():
            ops.ema_update(at.flat, as_.flat, float(beta))
            for n in at.names:
                done.add(id(at.params[n]))
            at._version_seen = None              # the bf16 shadows are stale: refreshed at the next forward
    for ps, pt in zip(current_model.parameters(), ma_model.parameters()):
        if id(pt) in done:
            continue
        if pt.is_cuda and pt.dtype == torch.float32 and pt.data.is_contiguous() and ps.data.is_contiguous():
            ops.ema_update(pt.data.view(-1), ps.detach().data.view(-1), float(beta))
        else:
            pt.data = pt.data * beta + (1.0 - beta) * ps.detach().data


class VTDINO(nn.Module):
    """Compute path of models/vtdino.py::VTDINO.  `encoder` is a `m3l_b200.vtt.VTT`; `dino_head` a partial of DINOHead
    without `in_dim` (as in the reference's hydra config)."""

    def __init__(self, encoder: nn.Module, dino_head: partial, optim_cfg=None, lr_scheduler_cfg=None, wd_scheduler_cfg=None,
                 online_probes=None, online_probes_lrs=(), local_mask_scale: Tuple[float, float] = (0.2, 0.8),
                 global_mask_scale: Tuple[float, float] = (0.2, 0.8), num_global_masks: int = 1, num_local_masks: int = 4,
                 min_keep_num_sensors: int = 4, allow_mask_overlap: bool = False,
                 moving_average_decay: Union[float, Tuple[float, ...]] = 0.99,
                 teacher_temp: Union[float, Tuple[float, ...]] = (0.04, 0.07), teacher_warmup_epochs: int = 10,
                 use_momentum=True, log_freq_reconstruction: int = 1000):
        super().__init__()
        if online_probes:
            raise M3LError("m3l_b200.VTDINO: online probes are outside the accelerated path")
        self.optim_partial, self.lr_scheduler_partial, self.wd_scheduler_partial = optim_cfg, lr_scheduler_cfg, wd_scheduler_cfg
        self.use_momentum = use_momentum
        self.global_mask_scale, self.local_mask_scale = global_mask_scale, local_mask_scale
        self.num_global_masks, self.num_local_masks = num_global_masks, num_local_masks
        self.min_keep = min_keep_num_sensors
        self.allow_mask_overlap = allow_mask_overlap
        self.generator = torch.Generator()
        self.step = -1
        dino_head = partial(dino_head, in_dim=encoder.embed_dim)
        self.student_encoder_dict, self.teacher_encoder_dict = dict(), dict()
        self.student_encoder_dict["backbone"] = encoder
        self.student_encoder_dict["dino_head"] = dino_head()
        self.student_encoder = nn.ModuleDict(self.student_encoder_dict)
        teacher = copy.deepcopy(encoder)
        if hasattr(teacher, "_arena"):
            teacher._arena = None                 # the copy binds its own flat arena on first use
        self.teacher_encoder_dict["backbone"] = teacher
        self.teacher_encoder_dict["dino_head"] = dino_head()
        self.teacher_encoder = nn.ModuleDict(self.teacher_encoder_dict)
        self.teacher_encoder.requires_grad_(False)
        self.dino_loss = DINOLoss(out_dim=self.student_encoder_dict["dino_head"].last_layer.out_features)
        self.patch_size = encoder.image_patch_height
        self.img_size = encoder.image_height
        self.in_chans = encoder.image_channels
        self.online_probes = []
        self.momentum_scheduler = None
        self.moving_average_decay = moving_average_decay if isinstance(moving_average_decay, float) else tuple(moving_average_decay)
        self.teacher_temp_scheduler = None
        self.teacher_temp = teacher_temp if isinstance(teacher_temp, float) else tuple(teacher_temp)
        self.current_teacher_temp = self.teacher_temp if isinstance(self.teacher_temp, float) else self.teacher_temp[0]
        self.teacher_warmup_epochs = teacher_warmup_epochs

    # ------------------------------------------------------------------ schedules (models/vtdino.py:486-503,548-565)
    def configure_schedules(self, num_iterations_per_epoch: int, num_epochs: int) -> None:
        total = int(num_epochs * num_iterations_per_epoch)
        if isinstance(self.moving_average_decay, tuple):
            m0, m1 = self.moving_average_decay
            self.momentum_scheduler = (m0 + i * (m1 - m0) / total for i in range(total + 1))
        if isinstance(self.teacher_temp, tuple):
            t0, t1 = self.teacher_temp
            warm = int(self.teacher_warmup_epochs * num_iterations_per_epoch)
            self.teacher_temp_scheduler = ((t0 + (t1 - t0) * min(i, warm) / max(warm, 1)) for i in range(total + 1))
            self.current_teacher_temp = t0

    def on_train_batch_end(self, outputs=None, batch=None, batch_idx=None, trainer_instance=None):
        self.current_teacher_temp = (next(self.teacher_temp_scheduler) if self.teacher_temp_scheduler is not None
                                     else (self.teacher_temp if isinstance(self.teacher_temp, float) else self.current_teacher_temp))
        if self.use_momentum:
            beta = next(self.momentum_scheduler) if self.momentum_scheduler is not None else self.moving_average_decay
            if isinstance(beta, tuple):
                beta = beta[0]
            with torch.no_grad():
                update_moving_average(self.teacher_encoder, self.student_encoder, beta)

    # ------------------------------------------------------------------ block masks (host side, models/vtdino.py:212-330)
    def _sample_block_size(self, height, width, scale):
        r = torch.rand(1, generator=self.generator).item()
        max_keep = int(height * width * (scale[0] + r * (scale[1] - scale[0])))
        h = min(int(round(math.sqrt(max_keep))), height)
        w = min(int(round(math.sqrt(max_keep))), width)
        return h, w

    def _sample_block_mask(self, height, width, b_size, acceptable_regions=None):
        h, w = b_size
        tries, timeout = 0, 20
        while True:
            top = int(torch.randint(0, height - h + 1, (1,), generator=self.generator))
            left = int(torch.randint(0, width - w + 1, (1,), generator=self.generator))
            mask = torch.zeros((height, width), dtype=torch.int32)
            mask[top:top + h, left:left + w] = 1
            if acceptable_regions is not None:
                for k in range(max(int(len(acceptable_regions) - tries), 0)):
                    mask *= acceptable_regions[k]
            idx = torch.nonzero(mask.flatten())
            if len(idx) > self.min_keep:
                break
            timeout -= 1
            if timeout == 0:
                tries, timeout = tries + 1, 20
        comp = torch.ones((height, width), dtype=torch.int32)
        comp[top:top + h, left:left + w] = 0
        return idx.squeeze(), comp

    def sample_masks(self, x):
        B, _, H, W = x.shape
        height, width = H // self.patch_size, W // self.patch_size
        lsize = self._sample_block_size(height, width, self.local_mask_scale)
        gsize = self._sample_block_size(height, width, self.global_mask_scale)
        loc, glo = [], []
        keep_l = keep_g = height * width
        for _ in range(B):
            ml, comps = [], []
            for _ in range(self.num_local_masks):
                m, c = self._sample_block_mask(height, width, lsize)
                ml.append(m); comps.append(c)
                keep_l = min(keep_l, len(m))
            loc.append(ml)
            regions = None if self.allow_mask_overlap else comps
            mg = []
            for _ in range(self.num_global_masks):
                m, _ = self._sample_block_mask(height, width, gsize, regions)
                mg.append(m)
                keep_g = min(keep_g, len(m))
            glo.append(mg)
        local_masks = [torch.stack([loc[b][p][:keep_l] for b in range(B)]).to(x.device) for p in range(self.num_local_masks)]
        global_masks = [torch.stack([glo[b][p][:keep_g] for b in range(B)]).to(x.device) for p in range(self.num_global_masks)]
        return global_masks, local_masks

    # ------------------------------------------------------------------ forward (models/vtdino.py:332-397)
    def forward(self, x: Dict[str, torch.Tensor], global_masks: List[torch.Tensor], local_masks: List[torch.Tensor]):
        assert global_masks is not None and local_masks is not None, "Masks are required for DINOModule during training"
        student, teacher = self.student_encoder_dict, self.teacher_encoder_dict
        sg = student["backbone"].forward_features(x, global_masks)["x_norm_regtokens"]          # ((p b), 1, c)
        pg, pl = len(global_masks), len(local_masks)
        B = sg.shape[0] // pg
        sg = sg.reshape(pg, B, -1).permute(1, 0, 2)                                              # (p b) 1 c -> b p c
        sl = student["backbone"].forward_features(x, local_masks)["x_norm_regtokens"]
        sl = sl.reshape(pl, B, -1).permute(1, 0, 2)
        s_cls = student["dino_head"](torch.cat([sg, sl], dim=-2))                                # b p c
        s_cls = s_cls.permute(1, 0, 2).unsqueeze(2)                                              # p b 1 c
        with torch.no_grad():
            tg = teacher["backbone"].forward_features(x, global_masks)["x_norm_regtokens"]
            t_cls = teacher["dino_head"](tg).detach()
            t_soft = self.dino_loss.softmax_center_teacher(t_cls, teacher_temp=self.current_teacher_temp)
            t_soft = t_soft.view(self.num_global_masks, -1, *t_cls.shape[1:])
            self.dino_loss.update_center(t_cls)
        return self.dino_loss(list(s_cls), list(t_soft))

    def training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0) -> Dict:
        self.step = self.step + 1
        self.generator.manual_seed(self.step)
        global_masks, local_masks = self.sample_masks(batch["image"])
        loss = self.forward(batch, global_masks, local_masks)
        return {"ssl_loss": loss.item(), "loss": loss, "online_probes_loss": 0.0}
