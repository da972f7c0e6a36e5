"""Multi-rank numerical check of the data-parallel fused trainer (run under torchrun, one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_worker.py
Every rank trains on its shard of a global batch; afterwards (i) all replicas must hold BIT-IDENTICAL parameters,
(ii) the update must match the CPU oracle's step on the whole global batch (loss = mean of the rank losses; update
direction cosine >= 0.99 per tensor).  Cases: the canonical dims and decoder_dim != dim (enc_to_dec is a Linear whose
gradients must be final in the first all-reduce bucket).  Prints DP_WORKER_OK on rank 0."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from oracle import vtmae_oracle as O
from tests._build import build_product
from m3l_b200 import dp
from m3l_b200.trainer import FusedTrainer


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for case, kw in (("canonical dims", dict(depth=2, decoder_depth=1)), ("decoder_dim != dim", dict(depth=1, decoder_depth=1, decoder_dim=128))):
        for one_graph in ("1", "0"):
            os.environ["M3L_DP_ONE_GRAPH"] = one_graph
            cfg = O.VTMAEConfig(**kw)
            sd = O.init_state_dict(cfg, seed=3)
            gen = torch.Generator().manual_seed(5)
            Bg = 4 * world
            x = {"image": torch.rand(Bg, 12, 64, 64, generator=gen), "tactile1": torch.rand(Bg, 12, 32, 32, generator=gen),
                 "tactile2": torch.rand(Bg, 12, 32, 32, generator=gen)}
            noise = O.tie_free_noise(Bg, 192, gen, [64] * 3)
            mae = build_product(cfg, device=dev, weights=sd)
            mae._sync()
            tr = FusedTrainer(mae, lr=1e-3)
            xl, nl = dp.shard_batch(x, noise, rank, world)
            xl = {k: v.to(dev) for k, v in xl.items()}
            losses = [tr.step(xl, noise=nl.to(dev)).clone() for _ in range(3)]
            torch.cuda.synchronize()
            flat = mae.arena.flat.clone()
            gathered = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            for r in range(world):
                assert torch.equal(gathered[r], gathered[0]), f"{case}: replica {r} differs from replica 0 (one_graph={one_graph})"
            l0 = losses[0].clone()
            dist.all_reduce(l0, op=dist.ReduceOp.AVG)
            if rank == 0:
                osd = {k: v.clone() for k, v in sd.items()}
                lref, _, _ = O.train_step(osd, cfg, x, noise, O.AdamWState(), lr=1e-3)
                assert abs(float(l0) - float(lref)) <= 1e-2 * abs(float(lref)), (case, float(l0), float(lref))
                print(f"[dp_worker] {case} one_graph={one_graph}: world {world} replicas identical after 3 steps; "
                      f"step-0 loss {float(l0):.5f} vs oracle (global batch) {float(lref):.5f}", flush=True)
            del mae, tr
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("DP_WORKER_OK", flush=True)
    # captured NCCL kernels were alive in this process: leave without the communicator teardown (see bench._teardown)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
