"""Flat parameter arena: the module's nn.Parameters become views into one fp32 buffer so that
(i) the fused clip+AdamW kernel and the NCCL gradient all-reduce run over contiguous memory, and
(ii) the bf16 shadow copies the tcgen05 GEMMs consume (plain and transposed) are refreshed with two
kernel launches.  Parameters stay ordinary nn.Parameters (state_dict / external optimizers keep
working: SURVEY.md §8b "ownership").
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import torch

from . import _lib, ops

_ALIGN = 64  # elements; keeps every tensor 256-byte aligned (float4 / TMA requirements)


def _round_up(x: int, a: int = _ALIGN) -> int:
    return (x + a - 1) // a * a


class ParamArena:
    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], device: torch.device,
                 late_names: Iterable[str] = ()):
        """named_params: (canonical name, parameter); duplicates (shared modules) are stored once.
        late_names: names placed at the end of the arena (parameters that usually get no gradient)."""
        seen = {}
        ordered: List[Tuple[str, torch.nn.Parameter]] = []
        for name, p in named_params:
            if id(p) in seen:
                continue
            seen[id(p)] = name
            ordered.append((name, p))
        late = set(late_names)
        ordered = [e for e in ordered if e[0] not in late] + [e for e in ordered if e[0] in late]
        self.device = device
        self.names: List[str] = []
        self.params: Dict[str, torch.nn.Parameter] = {}
        self.offset: Dict[str, int] = {}
        self.numel: Dict[str, int] = {}
        off = 0
        for name, p in ordered:
            self.names.append(name)
            self.params[name] = p
            self.offset[name] = off
            self.numel[name] = p.numel()
            off += _round_up(p.numel())
        self.total = off
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.flat_bf16 = torch.zeros(self.total, dtype=torch.bfloat16, device=device)
        # transposed bf16 copies of every 2-D weight (nn.Linear [out, in] -> [in, out])
        self.t_offset: Dict[str, int] = {}
        descs = []
        toff = 0
        for name, p in ordered:
            if p.dim() in (2, 4) and name.endswith(".weight") and min(p.shape[0], p.numel() // p.shape[0]) >= 8:
                # nn.Linear [out, in]; nn.Conv2d [cout, cin, kh, kw] is handled as [cout, cin*kh*kw]
                self.t_offset[name] = toff
                descs.append((self.offset[name], toff, p.shape[0], p.numel() // p.shape[0]))
                toff += _round_up(p.numel())
        self.flat_t = torch.zeros(max(toff, 1), dtype=torch.bfloat16, device=device)
        carr = (_lib.MatrixDesc * max(len(descs), 1))()
        for i, (so, do, r, c) in enumerate(descs):
            carr[i].src_offset, carr[i].dst_offset, carr[i].rows, carr[i].cols = so, do, r, c
        self.n_descs = len(descs)
        self.descs_dev = torch.frombuffer(bytearray(bytes(carr)), dtype=torch.uint8).to(device)
        self._version_seen = None
        self.bind()

    # ------------------------------------------------------------------------------ binding
    def bind(self) -> None:
        """Copies current parameter values into the arena and re-points .data at arena views."""
        with torch.no_grad():
            for name in self.names:
                p = self.params[name]
                v = self.view(self.flat, name)
                if p.data_ptr() != v.data_ptr():
                    v.copy_(p.data.to(device=self.device, dtype=torch.float32))
                    p.data = v
        self._version_seen = None

    def is_bound(self) -> bool:
        base = self.flat.data_ptr()
        return all(self.params[n].data_ptr() == base + 4 * self.offset[n] for n in self.names)

    def version(self) -> int:
        return sum(self.params[n]._version for n in self.names)

    def sync(self, force: bool = False) -> None:
        """Makes the bf16 shadows current (re-binding first if someone replaced parameter storage)."""
        if not self.is_bound():
            self.bind()
        v = self.version()
        if force or v != self._version_seen:
            self.refresh_shadows()
            self._version_seen = v

    def refresh_shadows(self) -> None:
        # the plain and the transposed bf16 copies only read the fp32 arena: two parallel graph branches
        from .engine import Branch
        br = Branch(self.device, enabled=bool(self.n_descs), index=2)
        with br:
            if self.n_descs:
                ops.transpose_cast_bf16(self.flat, self.flat_t, self.descs_dev, self.n_descs)
        ops.cast_bf16(self.flat, self.flat_bf16)
        br.join()

    # ------------------------------------------------------------------------------ accessors
    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o, n = self.offset[name], self.numel[name]
        return flat[o:o + n].view(self.params[name].shape)

    def f32(self, name: str) -> torch.Tensor:
        return self.view(self.flat, name)

    def bf(self, name: str) -> torch.Tensor:
        return self.view(self.flat_bf16, name)

    def bf_t(self, name: str) -> torch.Tensor:
        o, n = self.t_offset[name], self.numel[name]
        r = self.params[name].shape[0]
        return self.flat_t[o:o + n].view(n // r, r)

    def bf2d(self, name: str) -> torch.Tensor:
        """bf16 shadow of a weight as a 2-D [out, fan_in] matrix (conv weights flattened)."""
        o, n = self.offset[name], self.numel[name]
        return self.flat_bf16[o:o + n].view(self.params[name].shape[0], -1)

    def new_grad_buffer(self) -> torch.Tensor:
        return torch.zeros(self.total, dtype=torch.float32, device=self.device)

    def ranges(self, live_names: Iterable[str]) -> List[Tuple[int, int]]:
        """Merged [start, end) element ranges covering the given parameters (alignment gaps between
        adjacent live entries are included; they hold zeros)."""
        live = set(live_names)
        out: List[List[int]] = []
        prev_live_end_idx = None
        for i, name in enumerate(self.names):
            if name not in live:
                continue
            s = self.offset[name]
            e = s + _round_up(self.numel[name])
            if out and prev_live_end_idx == i - 1:
                out[-1][1] = e
            else:
                out.append([s, e])
            prev_live_end_idx = i
        return [(s, e) for s, e in out]
