"""Synchronous data parallelism for the MSDN/DCNF step: one process per GPU, full weight replica per
rank, gradient sum-allreduce (NCCL over NVLink 5 / NVSwitch) of flat arena buckets on a dedicated
communication stream, overlapped with the rest of the backward pass.

Replaces the asynchronous parameter-server replication of the reference (src/ann3depth.py:78-92,
`replica_device_setter` + gRPC variable pulls / gradient pushes every `session.run`).  Semantics
change: the reference applies stale per-worker gradients; here n ranks x batch B behave like one
rank x batch n*B (gradients averaged: sum-allreduce, 1/n folded into the optimizer's grad_scale).
"""
from __future__ import annotations

import torch

from . import ops


class DataParallel:
    def __init__(self, ctx: ops.Context, rank: int, world: int, unique_id: bytes, nccl_path: str | None = None):
        self.ctx, self.rank, self.world = ctx, rank, world
        ctx.comm_init(unique_id, rank, world, nccl_path)
        self.stream = torch.cuda.Stream(device=ctx.device)
        self._done = []
        self.bytes_per_step = 0

    @staticmethod
    def bucket_range(net, name):
        a = net.arena
        if name in a.groups:
            return a.group_range(name)
        if name == "coarse_conv":
            return a.group_range("CoarseConv")
        if name == "fine":
            lo_a, hi_a = a.group_range("FineA")
            lo_b, hi_b = a.group_range("FineB")
            return min(lo_a, lo_b), max(hi_a, hi_b)
        if name == "all":
            return 0, a.total
        # a single layer: kernel + bias segments are adjacent
        keys = [k for k in a.specs if k.startswith(name + "/") or ("/" + name + "/") in k]
        lo = min(a.specs[k].offset for k in keys)
        hi = max(a.specs[k].offset + a.specs[k].size for k in keys)
        return lo, hi

    def bucket_ready(self, net, name):
        """Called right after the wgrad kernels of a bucket were enqueued on the compute stream."""
        lo, hi = self.bucket_range(net, name)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.ctx.allreduce_sum(net.arena.g[lo:hi])
            done = torch.cuda.Event()
            done.record(self.stream)
        self._done.append(done)
        self.bytes_per_step += (hi - lo) * 4

    def wait_all(self, net):
        cur = torch.cuda.current_stream()
        for e in self._done:
            cur.wait_event(e)
        self._done = []
