"""Host-side input contract of the hot path: `vt_load` (obs dict -> model input dict), same
semantics as /root/reference/utils/pretrain_utils.py:7-57 — image NHWC(3F) -> NCHW scaled to [0,1],
tactile (B, 3F*sensors, h, w) de-interleaved per sensor and mapped from [-1,1] to [0,1]."""
from __future__ import annotations

import numpy as np
import torch


def vt_load(x, image_normalization=(0, 1), tactile_normalization=(-1, 1), squeeze=False, frame_stack=1):
    if isinstance(x, str):
        x = np.load(x, allow_pickle=True).item()
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
        # channels {i * per_frame + 3 s + c : i < frame_stack, c < 3} of sensor s (pretrain_utils.py:36-49), taken as a
        # view + slice (no index tensor: the rollout path replays this inside a CUDA graph)
        b, _, h, w = tac.shape
        frames = tac.reshape(b, frame_stack, per_frame, h, w)
        for s in range(per_frame // 3):
            x[f"tactile{s + 1}"] = (frames[:, :, 3 * s:3 * s + 3].reshape(b, 3 * frame_stack, h, w) - lo) / (hi - lo)
        del x["tactile"]
    if squeeze:
        for key in x:
            x[key] = x[key].squeeze()
    return x
