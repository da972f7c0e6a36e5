"""Data-parallel plumbing for the VTMAE step (one process per GPU, torch.distributed).

The path shards by independent samples: rank r owns samples [r*B/W, (r+1)*B/W) and
the matching rows of the mask noise; replicas hold identical weights and optimizer state.  The only
exchange is the gradient all-reduce (AVG) over the flat gradient arena, issued in two buckets —
heads+decoder first (their gradients are final first in backward), encoder+embeddings second.
These helpers are device-agnostic so the host logic is testable with gloo on CPU.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

DECODER_SIDE_PREFIXES = ("to_pixels", "to_tactiles", "decoder.", "mask_token", "decoder_modality_embedding",
                         "enc_to_dec", "decoder_pos_emb")


def shard_batch(x: Dict[str, torch.Tensor], noise: torch.Tensor, rank: int, world: int):
    """Rank-local slice of a global batch (global batch size must divide by world)."""
    B = noise.shape[0]
    if B % world:
        raise ValueError(f"global batch {B} does not divide by world size {world}")
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    return {k: v[lo:hi] for k, v in x.items()}, noise[lo:hi]


def split_buckets(live_names: Sequence[str]) -> Tuple[List[str], List[str]]:
    """(decoder-side names, encoder-side names) in arena order."""
    dec = [k for k in live_names if k.startswith(DECODER_SIDE_PREFIXES)]
    enc = [k for k in live_names if not k.startswith(DECODER_SIDE_PREFIXES)]
    return dec, enc


ENCODER_STACK_PREFIX = "encoder.transformer."


def split_three(live_names: Sequence[str]) -> Tuple[List[str], List[str], List[str]]:
    """(heads + decoder, encoder transformer, token embeddings) in the order their gradients become final in the
    backward pass: three all-reduce buckets, each overlapped with the next phase of the backward."""
    dec, enc = split_buckets(live_names)
    stack = [k for k in enc if k.startswith(ENCODER_STACK_PREFIX)]
    emb = [k for k in enc if not k.startswith(ENCODER_STACK_PREFIX)]
    return dec, stack, emb


def allreduce_ranges(flat: torch.Tensor, ranges: Sequence[Tuple[int, int]], group=None, average: bool = True):
    """All-reduces flat[s:e] for every range (in place).  AVG where the backend has it (NCCL), else
    SUM followed by a scale (gloo)."""
    world = dist.get_world_size(group)
    for s, e in ranges:
        view = flat[s:e]
        if average and dist.get_backend(group) == "nccl":
            dist.all_reduce(view, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
            if average:
                view.div_(world)
