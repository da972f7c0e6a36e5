"""TEST INFRASTRUCTURE (oracle) — restatement of `positional-encodings==6.0.1`
`PositionalEncoding2D` (pinned at /root/reference/requirements.txt:107; not in the reference tree),
called at /root/reference/models/pretrain_models.py:120-140.

PARITY UNPINNED: restated from the package's published algorithm (interleaved sin/cos of
pos * 10000^(-2i/ch), first half of the channels for the row index, second half for the column).
"""
import numpy as np
import torch
from torch import nn


def _interleaved_sin_cos(angles):
    # [..., f] -> [..., 2f] laid out sin f0, cos f0, sin f1, cos f1, ...
    return torch.stack((angles.sin(), angles.cos()), dim=-1).flatten(-2, -1)


class PositionalEncoding2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.org_channels = channels
        channels = int(np.ceil(channels / 4) * 2)
        self.channels = channels
        inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
        self.register_buffer("inv_freq", inv_freq)
        self.register_buffer("cached_penc", None, persistent=False)

    def forward(self, tensor):
        if tensor.dim() != 4:
            raise RuntimeError("The input tensor has to be 4d!")
        if self.cached_penc is not None and self.cached_penc.shape == tensor.shape:
            return self.cached_penc
        self.cached_penc = None
        batch, nx, ny, orig_ch = tensor.shape
        pos_x = torch.arange(nx, device=tensor.device, dtype=self.inv_freq.dtype)
        pos_y = torch.arange(ny, device=tensor.device, dtype=self.inv_freq.dtype)
        emb_x = _interleaved_sin_cos(torch.einsum("i,j->ij", pos_x, self.inv_freq)).unsqueeze(1)
        emb_y = _interleaved_sin_cos(torch.einsum("i,j->ij", pos_y, self.inv_freq))
        emb = torch.zeros((nx, ny, self.channels * 2), device=tensor.device, dtype=tensor.dtype)
        emb[:, :, : self.channels] = emb_x
        emb[:, :, self.channels : 2 * self.channels] = emb_y
        self.cached_penc = emb[None, :, :, :orig_ch].repeat(tensor.shape[0], 1, 1, 1)
        return self.cached_penc
