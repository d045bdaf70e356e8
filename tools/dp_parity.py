#!/usr/bin/env python
"""Multi-GPU check (torchrun): one general-Adam (beta2 = 0.999) phase-1 step under the two data-parallel dense
paths -- activation gather (dp.dense_gather_adam) and gradient reduce-scatter (dp.sharded_adam) -- must produce the
same weight update on every rank, and all ranks must hold identical bf16 weights afterwards."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ann3depth_b200 import models, ops  # noqa: E402
from ann3depth_b200.dp import DataParallel  # noqa: E402
from ann3depth_b200.init import glorot_params  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("gloo")
    ctx = models.get_context(local)
    saved = os.dup(1); os.dup2(2, 1)
    ids = [ops.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = DataParallel(ctx, rank, world, ids[0])
    sys.stdout.flush(); os.dup2(saved, 1)
    images, depths = bench.synthetic_batch(rank, torch)
    res = {}
    for mode in ("1", "0"):
        os.environ["A3D_DP_GATHER"] = mode
        op = models.msdn(images.to(dev), depths.to(dev), train=True, comm=comm, beta2=0.999)
        p = glorot_params(seed=1)
        p["coarse/dense/dense_1/bias"] = p["coarse/dense/dense_1/bias"] + 1.0
        op.net.load_params(p)
        w0 = op.net.arena.wb.float().clone()
        op.run(use_graph=False)
        torch.cuda.synchronize()
        res[mode] = (op.net.arena.wb.float() - w0).cpu()
        # every rank must hold the same mirror
        mine = op.net.arena.wb.float().cpu()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(mine, ref))
        flags = [None] * world
        dist.all_gather_object(flags, same)
        if rank == 0:
            print(f"gather={mode}: all ranks hold identical bf16 weights: {all(flags)}", flush=True)
        del op
    a, b = res["1"].double(), res["0"].double()
    cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
    if rank == 0:
        print(f"update cosine (activation gather vs reduce-scatter): {cos:.6f}; |dw| {float(a.norm()):.4f} vs {float(b.norm()):.4f}",
              flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if cos > 0.99 else 1)


if __name__ == "__main__":
    main()
