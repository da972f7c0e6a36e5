"""Input contract of the hot path: `vt_load` (obs dict -> model input dict), same semantics as
/root/reference/utils/pretrain_utils.py:7-57 — image NHWC(3F) -> NCHW scaled to [0,1], tactile
(B, 3F*sensors, h, w) de-interleaved per sensor and mapped from [-1,1] to [0,1].

Two forms:
  * eager (default, the reference's behaviour): returns fp32 NCHW tensors.  CUDA observations are converted by the
    `m3l_vt_load` kernel, CPU / numpy observations by the same torch ops the reference uses.
  * lazy (`vt_load(..., lazy=True)`, CUDA observations): returns `RawMap` views — the observation tensors are left
    as the rollout buffer holds them (image [B, H, W, 3F] or [B, F, H, W, 3], fp32 or uint8 frames; tactile
    [B, 3F*sensors, h, w] or [B, F, 3*sensors, h, w]) and the layout change, the per-sensor channel de-interleave
    and the normalisation happen INSIDE the patch-gather kernels of the train step / encoder pass
    (`m3l_patch_source.layout == 1`, include/m3l_b200.h).  VTMAE.forward / get_embeddings / train_step and
    MAEExtractor accept either form; MAEExtractor and train_iterations use the lazy form themselves.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch


class RawMap:
    """A model input map [B, C, H, W] that still lives inside a raw observation tensor.

    `t` is the observation tensor (contiguous, fp32 or uint8, on a CUDA device); the remaining fields say where
    element (b, c, y, x), c = f * cg + ch, sits in it (element strides) and how it is normalised:
    (raw - lo) / span in fp32, uint8 frames as raw / 255 first."""

    __slots__ = ("t", "shape", "offset", "sb", "cg", "sf", "sch", "sy", "sx", "lo", "span")

    def __init__(self, t, shape, offset, sb, cg, sf, sch, sy, sx, lo, span, strided_view=False):
        # strided_view: `t` is itself a strided view whose data_ptr() is the origin (offset 0); such maps cannot serve
        # as static graph inputs (clone_inputs) and are only built by callers that read them immediately
        assert (strided_view or t.is_contiguous()) and t.dtype in (torch.float32, torch.uint8), \
            "raw observations must be contiguous fp32 / uint8"
        self.t, self.shape, self.offset = t, tuple(shape), int(offset)
        self.sb, self.cg, self.sf, self.sch, self.sy, self.sx = int(sb), int(cg), int(sf), int(sch), int(sy), int(sx)
        self.lo, self.span = float(lo), float(span)

    # --- the little of the tensor interface the callers of the reference's vt_load use on its results
    @property
    def is_cuda(self):
        return self.t.is_cuda

    @property
    def device(self):
        return self.t.device

    @property
    def dtype(self):
        return torch.float32

    def dim(self):
        return 4

    def _with(self, t):
        return RawMap(t, self.shape, self.offset, self.sb, self.cg, self.sf, self.sch, self.sy, self.sx, self.lo, self.span)

    def to(self, *args, **kwargs):
        kwargs.pop("dtype", None)
        args = tuple(a for a in args if not isinstance(a, torch.dtype))
        t = self.t.to(*args, **kwargs)
        return self if t is self.t else self._with(t.contiguous())

    def cuda(self, *a, **k):
        return self.to("cuda")

    def detach(self):
        return self

    def data_ptr(self):
        return self.t.data_ptr() + self.offset * self.t.element_size()

    def materialize(self) -> torch.Tensor:
        """fp32 [B, C, H, W] contiguous (the eager vt_load result) through the m3l_vt_load kernel."""
        from . import ops
        return ops.vt_load_map(self)


def _image_view(img: torch.Tensor, frame_stack: int, lo, hi) -> RawMap:
    if img.dim() == 5:                                   # [B, F, H, W, 3]  (pretrain_models.py:823-824 reshapes it)
        B, F, H, W, c3 = img.shape
        assert F == frame_stack and c3 == 3, f"image {tuple(img.shape)} is not [B, {frame_stack}, H, W, 3]"
        return RawMap(img, (B, 3 * F, H, W), 0, F * H * W * 3, 3, H * W * 3, 1, W * 3, 3, lo, hi - lo)
    B, H, W, C = img.shape                               # [B, H, W, 3F]
    assert C == 3 * frame_stack
    return RawMap(img, (B, C, H, W), 0, H * W * C, C, 0, 1, W * C, C, lo, hi - lo)     # one channel group: (x, c) contiguous


def _tactile_views(tac: torch.Tensor, frame_stack: int, lo, hi) -> Dict[str, RawMap]:
    if tac.dim() == 5:                                   # [B, F, 3*sensors, h, w] == [B, F*3*sensors, h, w] in memory
        B, F, per_frame, h, w = tac.shape
        assert F == frame_stack
    else:
        B, ch, h, w = tac.shape
        assert ch in (3 * frame_stack, 6 * frame_stack, 12 * frame_stack)
        per_frame = ch // frame_stack
    out = {}
    for s in range(per_frame // 3):                      # channels {i*per_frame + 3s + c} of sensor s (pretrain_utils.py:36-49)
        out[f"tactile{s + 1}"] = RawMap(tac, (B, 3 * frame_stack, h, w), 3 * s * h * w, frame_stack * per_frame * h * w, 3,
                                        per_frame * h * w, h * w, w, 1, lo, hi - lo)
    return out


def vt_load_lazy(x, image_normalization=(0, 1), tactile_normalization=(-1, 1), frame_stack=1, device=None) -> Dict[str, RawMap]:
    """Raw observation dict -> {image, tactile1, ...} of RawMap views (no data movement besides an H2D copy of
    host observations).  Accepts the 4-D forms the reference hands to vt_load and the 5-D frame-stacked forms."""
    out = {}
    for key in ("image", "tactile"):
        if key not in x:
            continue
        t = torch.as_tensor(x[key])
        if device is not None:
            t = t.to(device, non_blocking=True)
        if t.dtype not in (torch.float32, torch.uint8):
            t = t.to(torch.float32)
        if t.dim() == 3:
            t = t[None]
        t = t.contiguous()
        if key == "image":
            out["image"] = _image_view(t, frame_stack, *image_normalization)
        else:
            assert t.dtype == torch.float32, "tactile observations are fp32"
            out.update(_tactile_views(t, frame_stack, *tactile_normalization))
    return out


def vt_load(x, image_normalization=(0, 1), tactile_normalization=(-1, 1), squeeze=False, frame_stack=1, lazy=False):
    if isinstance(x, str):
        x = np.load(x, allow_pickle=True).item()
    if lazy:
        assert not squeeze
        return vt_load_lazy(x, image_normalization, tactile_normalization, frame_stack)
    on_gpu = any(isinstance(x.get(k), torch.Tensor) and x[k].is_cuda for k in ("image", "tactile"))
    if on_gpu:
        views = vt_load_lazy(x, image_normalization, tactile_normalization, frame_stack)
        x.pop("tactile", None)
        for k, v in views.items():
            x[k] = v.materialize()
    else:
        for key in ("image", "tactile"):
            if key in x and len(x[key].shape) == 3:
                x[key] = x[key][None]
        if "image" in x:
            assert x["image"].shape[-1] == 3 * frame_stack
            img = torch.as_tensor(x["image"]).to(torch.float32).permute(0, 3, 1, 2)
            lo, hi = image_normalization
            x["image"] = (img - lo) / (hi - lo)
        if "tactile" in x:
            ch = x["tactile"].shape[1]
            assert ch in (3 * frame_stack, 6 * frame_stack, 12 * frame_stack)
            per_frame = ch // frame_stack
            tac = torch.as_tensor(x["tactile"]).to(torch.float32)
            lo, hi = tactile_normalization
            b, _, h, w = tac.shape
            frames = tac.reshape(b, frame_stack, per_frame, h, w)
            for s in range(per_frame // 3):
                x[f"tactile{s + 1}"] = (frames[:, :, 3 * s:3 * s + 3].reshape(b, 3 * frame_stack, h, w) - lo) / (hi - lo)
            del x["tactile"]
    if squeeze:
        for key in x:
            x[key] = x[key].squeeze()
    return x


# --------------------------------------------------------------------------------------------
# static input buffers of captured CUDA graphs (plain tensors or RawMap views; the tactile sensors of one raw
# observation share ONE buffer)
# --------------------------------------------------------------------------------------------
def clone_inputs(xs: dict) -> dict:
    memo, out = {}, {}
    for k, v in xs.items():
        if isinstance(v, RawMap):
            t = memo.get(id(v.t))
            if t is None:
                t = memo[id(v.t)] = v.t.clone()
            out[k] = v._with(t)
        else:
            out[k] = v.clone()
    return out


def copy_inputs(dst: dict, src: dict) -> None:
    done = set()
    for k, v in src.items():
        d = dst[k]
        if isinstance(d, RawMap):
            if not isinstance(v, RawMap) or v.t.shape != d.t.shape or v.t.dtype != d.t.dtype:
                raise ValueError(f"input '{k}': raw observation layout changed between calls of the same shape key")
            if id(d.t) not in done:
                d.t.copy_(v.t, non_blocking=True)
                done.add(id(d.t))
        else:
            d.copy_(v, non_blocking=True)


def input_signature(xs: dict):
    """Part of a graph-cache key: distinguishes tensor inputs from raw views (and their layouts / dtypes)."""
    sig = []
    for k, v in sorted(xs.items()):
        if isinstance(v, RawMap):
            sig.append((k, "raw", tuple(v.t.shape), v.t.dtype, v.offset))
        else:
            sig.append((k, "map"))
    return tuple(sig)
