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
wsum = op.net.arena.wb.double().sum().item()
lst = [None] * world
dist.all_gather_object(lst, wsum)
say(f"bf16 weight mirror checksums equal across ranks: {lst}")
# equivalence: n ranks fed the SAME batch must take the single-GPU step (gradients are averaged)
gs = torch.Generator().manual_seed(7)
im2 = torch.rand(4, 480, 640, 3, generator=gs).cuda()
dp2 = (torch.rand(4, 55, 73, 1, generator=gs) * 0.95 + 0.05).cuda()
mask2 = (torch.rand(4, 4096, generator=gs) < 0.5).cuda()
pp = glorot_params(seed=1)
pp["coarse/dense/dense_1/bias"] += 1.0
outs = []
for cm in (None, comm):
    o2 = models.msdn(im2, dp2, train=True, comm=cm, beta2=0.999)
    o2.net.load_params(pp)
    o2.net.set_dropout_mask(mask2)
    w0 = o2.net.arena.wb.float().clone()
    o2.run(use_graph=False)
    torch.cuda.synchronize()
    if cm is not None:
        cm.gather_master(o2.net)
        torch.cuda.synchronize()
    outs.append((o2.net.arena.wb.float() - w0, o2.net.arena.m.clone(), o2.net.arena.w.clone()))
(d1, m1, w1), (d2, m2, w2) = outs
lo, hi = o2.net.arena.group_range("CoarseDense")[0], o2.net.arena.group_range("CoarseConv")[1]
cosd = float((d1[lo:hi].double() @ d2[lo:hi].double()) / (d1[lo:hi].double().norm() * d2[lo:hi].double().norm()))
cosm = float((m1[lo:hi].double() @ m2[lo:hi].double()) / (m1[lo:hi].double().norm() * m2[lo:hi].double().norm()))
say(f"DP(same batch) vs single GPU: update cosine {cosd:.6f}  m cosine {cosm:.6f}  max|w diff| {float((w1[lo:hi]-w2[lo:hi]).abs().max()):.3e}")
assert cosd > 0.98 and cosm > 0.999
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
