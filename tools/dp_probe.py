"""Stage-by-stage probe of the data-parallel path (run under torchrun with a short timeout)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

t00 = time.time()


def say(msg):
    print(f"[rank {os.environ.get('RANK')}] +{time.time() - t00:6.1f}s {msg}", file=sys.stderr, flush=True)


rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
say("gloo up")
from ann3depth_b200 import models, ops
from ann3depth_b200.dp import DataParallel
from ann3depth_b200.init import glorot_params
ctx = models.get_context(lr)
ids = [ops.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
say("id broadcast")
comm = DataParallel(ctx, rank, world, ids[0])
say("nccl comm up")
t = torch.full((1 << 20,), float(rank + 1), device=f"cuda:{lr}")
ctx.allreduce_sum(t)
torch.cuda.synchronize()
say(f"allreduce ok: {float(t[0])} (expect {world * (world + 1) / 2})")
g = torch.Generator().manual_seed(100 + rank)
images = torch.rand(32, 480, 640, 3, generator=g).cuda()
depths = (torch.rand(32, 55, 73, 1, generator=g) * 0.95 + 0.05).cuda()
op = models.msdn(images, depths, train=True, comm=comm)
op.net.load_params(glorot_params(seed=1))
say("model built")
op.run(use_graph=False)
torch.cuda.synchronize()
say("eager DP step ok")
gsum = op.net.arena.g[:4096].double().sum().item()
lst = [None] * world
dist.all_gather_object(lst, gsum)
say(f"grad checksums equal across ranks: {lst}")
if os.environ.get("A3D_PROBE_GRAPH", "1") == "1":
    op.run(use_graph=True)
    torch.cuda.synchronize()
    say("graph capture + replay ok")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        op.run()
    e1.record()
    torch.cuda.synchronize()
    say(f"10 graph steps: {e0.elapsed_time(e1) / 10:.3f} ms/step")
dist.barrier()
say("done")
sys.stderr.flush()
os._exit(0)
