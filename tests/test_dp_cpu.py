"""CPU tests (gloo, world_size 2) of the data-parallel protocol and host-side layout logic.

The GPU path sum-allreduces flat gradient buckets with NCCL and folds 1/n into the optimizer
(ann3depth_b200/dp.py).  Here the same protocol runs over gloo with the CPU oracle as the per-rank
model: n ranks x batch B must equal one rank x batch n*B."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ann3depth_b200.dp import DataParallel
from ann3depth_b200.params import Arena, msdn_specs, pack, unpack
from oracle import msdn as OM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(n, 480, 640, 3, generator=g)
    depths = torch.rand(n, 55, 73, 1, generator=g) * 0.95 + 0.05
    mask = (torch.rand(n, 4096, generator=g) < 0.5).float()
    return images, depths, mask


def _params():
    p = OM.init_params(1, torch.float32, bias_range=0.05)
    p["coarse/dense/dense_1/bias"] += 1.0
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    images, depths, mask = _batch(world)
    p = _params()
    sl = slice(rank, rank + 1)                                    # shard: one sample per rank
    g, _ = OM.grads(p, images[sl], depths[sl], mask[sl], "coarse")
    # flat bucket in arena order, sum-allreduce, then the optimizer's grad_scale = 1/n
    names = [n for n in g]
    flat = torch.cat([g[n].reshape(-1) for n in names])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat *= 1.0 / world
    if rank == 0:
        torch.save({"names": names, "flat": flat}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dp_mean_of_rank_gradients_equals_big_batch(tmp_path):
    world, port = 2, 29531
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    images, depths, mask = _batch(world)
    ref, _ = OM.grads(_params(), images, depths, mask, "coarse")
    flat_ref = torch.cat([ref[n].reshape(-1) for n in got["names"]])
    cos = float(got["flat"] @ flat_ref / (got["flat"].norm() * flat_ref.norm()))
    assert cos > 0.99999
    assert float((got["flat"] - flat_ref).abs().max() / flat_ref.abs().max()) < 1e-3


class _Net:
    def __init__(self):
        self.arena = Arena(msdn_specs(), "cpu", with_adam=False)


def test_arena_layout_and_buckets():
    net = _Net()
    a = net.arena
    assert a.num_real_params() == 70877171                      # KA2
    # every segment 16-byte aligned in both the f32 and the bf16 view
    for s in a.specs.values():
        assert s.offset % 64 == 0 and s.size % 64 == 0
    # optimizer groups are contiguous, disjoint, and cover the arena in backward order
    order = ["CoarseDense", "CoarseConv", "FineA", "FineB"]
    pos = 0
    for g in order:
        lo, hi = a.group_range(g)
        assert lo == pos
        pos = hi
    assert pos == a.total
    br = DataParallel.bucket_range
    d1, d0, cc = br(net, "dense_1"), br(net, "dense_0"), br(net, "coarse_conv")
    assert d1[0] == 0 and d1[1] == d0[0] and d0[1] == cc[0]      # buckets become ready front to back
    assert br(net, "CoarseDense") == (d1[0], d0[1])
    fa, fb = a.group_range("FineA"), a.group_range("FineB")
    assert br(net, "fine") == (fa[0], fb[1])


def test_pack_unpack_roundtrip_and_masks():
    a = _Net().arena
    g = torch.Generator().manual_seed(0)
    for name in ("coarse/conv/conv2d_0/kernel", "fine/first/conv2d/kernel", "fine/first/conv2d/bias",
                 "coarse/dense/dense_1/kernel", "coarse/conv/conv2d_3/kernel"):
        s = a.specs[name]
        t = torch.rand(s.tf_shape, generator=g)
        assert torch.equal(unpack(s, pack(s, t)), t)
    m = a.masks["coarse/conv/conv2d_0/kernel"].view(96, 3, 3, 64)        # space-to-depth(4) packing of 11x11x3
    assert int(m.sum()) == 96 * 11 * 11 * 3 and not bool(m[..., 48:].any())
    assert not bool(m[:, 2, :, 36:48].any())                              # filter row 11 (= 4*2+3) does not exist
    m = a.masks["fine/first/conv2d/kernel"].view(64, 5, 5, 16)
    assert int(m.sum()) == 63 * 9 * 9 * 3 and not bool(m[63].any())
    assert "coarse/conv/conv2d_2/kernel" not in a.masks
    m = a.masks["coarse/conv/conv2d_1/kernel"].view(256, 5, 5, 128)      # input channels stored 96 -> 128
    assert int(m.sum()) == 256 * 5 * 5 * 96 and not bool(m[..., 96:].any())


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, K, N = 4, 64, 48
    g = torch.Generator().manual_seed(10 + rank)
    x, dy = torch.randn(B, K, generator=g), torch.randn(B, N, generator=g)
    # activation-gather protocol of dp.dense_gather_adam: gather the small matrices, update own rows only
    xs, dys = [torch.empty_like(x) for _ in range(world)], [torch.empty_like(dy) for _ in range(world)]
    dist.all_gather(xs, x)
    dist.all_gather(dys, dy)
    x_all, dy_all = torch.cat(xs), torch.cat(dys)
    r = N // world
    mine = (dy_all.t() @ x_all)[rank * r:(rank + 1) * r] / world
    # reference protocol: per-rank gradient, sum-allreduce, 1/n
    full = dy.t() @ x
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    full /= world
    rows = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine)                                   # = the all-gather of the updated weight rows
    if rank == 0:
        torch.save({"gathered": torch.cat(rows), "allreduced": full}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dense_activation_gather_equals_gradient_allreduce(tmp_path):
    """dW = dy^T x has rank B: all-gathering x and dy and updating one's own rows is the same update as the
    gradient sum-allreduce it replaces (dp.dense_gather_adam, a3d_dense_wgrad_adam_rows)."""
    world, port = 2, 29532
    out = str(tmp_path / "gather.pt")
    mp.spawn(_gather_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    assert torch.allclose(got["gathered"], got["allreduced"], atol=1e-5)


def test_dense_gather_eligibility():
    net = _Net()
    dp = DataParallel.__new__(DataParallel)
    for world, ok in ((2, True), (4, True), (8, True), (3, False)):
        dp.world = world
        for k in ("coarse/dense/dense_1/kernel", "coarse/dense/dense_0/kernel"):
            assert dp.can_gather_dense(net, k, 32) is ok, (world, k)
    dp.world = 8
    assert not dp.can_gather_dense(net, "coarse/dense/dense_0/kernel", 64)       # gathered batch 512 > 256
    s = net.arena.specs["coarse/dense/dense_1/kernel"]
    assert s.packed_shape == (4096, 4096) and s.tf_shape == (4096, 4070)          # rows padded for equal slices
    t = torch.rand(4096, 4070)
    assert torch.equal(unpack(s, pack(s, t)), t) and float(pack(s, t)[4070:].abs().max()) == 0.0


def test_bench_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
