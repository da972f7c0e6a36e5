"""CPU, world_size 2, gloo: the data-parallel host logic of the train step — batch sharding,
gradient buckets and the averaged all-reduce — reproduces the single-process global-batch gradient
of the oracle (global loss = mean of the per-rank losses when the local batches are equal)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from m3l_b200 import dp
from oracle import vtmae_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flatten(grads, keys):
    return torch.cat([grads[k].flatten() for k in keys])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = O.VTMAEConfig(depth=1, decoder_depth=1)
    sd = O.init_state_dict(cfg, seed=0)
    gen = torch.Generator().manual_seed(3)
    B = 4
    x = {"image": torch.rand(B, 12, 64, 64, generator=gen), "tactile1": torch.rand(B, 12, 32, 32, generator=gen),
         "tactile2": torch.rand(B, 12, 32, 32, generator=gen)}
    noise = O.tie_free_noise(B, 192, gen, [64, 64, 64])
    xs, ns = dp.shard_batch(x, noise, rank, world)
    keys = O.param_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    loss = O.vtmae_forward(sd, cfg, xs, ns)
    loss.backward()
    live = [k for k in keys if sd[k].grad is not None]
    dec, enc = dp.split_buckets(live)
    assert set(dec) | set(enc) == set(live) and not (set(dec) & set(enc))
    assert all(k.startswith(dp.DECODER_SIDE_PREFIXES) for k in dec)
    order = dec + enc                                      # arena order: decoder side first
    flat = _flatten({k: sd[k].grad for k in live}, order)
    n_dec = sum(sd[k].numel() for k in dec)
    dp.allreduce_ranges(flat, [(0, n_dec)])                # bucket 1 (overlaps the encoder backward on GPU)
    dp.allreduce_ranges(flat, [(n_dec, flat.numel())])     # bucket 2
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    if rank == 0:
        # plain numpy through the queue: a torch tensor would travel as a shared-memory file descriptor that the
        # parent can only receive while this process is still alive (EOFError in recvfds otherwise)
        out.put((order, flat.numpy().copy(), (lt / world).numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_average_equals_global_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    order, flat, loss = out.get(timeout=240)
    flat, loss = torch.from_numpy(flat), torch.from_numpy(loss)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, global batch
    cfg = O.VTMAEConfig(depth=1, decoder_depth=1)
    sd = O.init_state_dict(cfg, seed=0)
    gen = torch.Generator().manual_seed(3)
    B = 4
    x = {"image": torch.rand(B, 12, 64, 64, generator=gen), "tactile1": torch.rand(B, 12, 32, 32, generator=gen),
         "tactile2": torch.rand(B, 12, 32, 32, generator=gen)}
    noise = O.tie_free_noise(B, 192, gen, [64, 64, 64])
    for k in O.param_keys(sd):
        sd[k].requires_grad_(True)
    ref = O.vtmae_forward(sd, cfg, x, noise)
    ref.backward()
    assert torch.allclose(loss, ref.detach(), rtol=1e-5)
    ref_flat = _flatten({k: sd[k].grad for k in order}, order)
    assert torch.allclose(flat, ref_flat, rtol=1e-4, atol=1e-7)


def test_shard_batch_rejects_ragged_global_batch():
    x = {"image": torch.zeros(5, 3, 8, 8)}
    with pytest.raises(ValueError):
        dp.shard_batch(x, torch.zeros(5, 4), 0, 2)
    xs, ns = dp.shard_batch({"image": torch.arange(8).reshape(8, 1)}, torch.arange(8).reshape(8, 1), 1, 4)
    assert xs["image"].flatten().tolist() == [2, 3] and ns.flatten().tolist() == [2, 3]
