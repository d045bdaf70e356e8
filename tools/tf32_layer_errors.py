"""Per-layer error of the TF32 mode against the float64 oracle (which layer the 1e-4 .. 4e-4 comes from): a debugging probe."""
import sys, torch
sys.path.insert(0, "/root/repo")
from oracle import msdn as OM, tf1_ops as T
from ann3depth_b200 import models
B=4
g = torch.Generator().manual_seed(0)
images = torch.rand(B, 480, 640, 3, generator=g); depths = torch.rand(B, 55, 73, 1, generator=g) * 0.95 + 0.05
p = OM.init_params(1, torch.float32, bias_range=0.05); p["coarse/dense/dense_1/bias"] += 1.0; p["fine/third/bias"] += 1.0
op = models.msdn(images.cuda(), depths.cuda(), train=False, dtype="tf32"); op.net.load_params(p); op.run(); torch.cuda.synchronize()
n = op.net
p64 = {k: v.double() for k, v in p.items()}
im, dp = OM.preprocess(images.double(), depths.double())
def l2(a, b): a=a.double().cpu(); return float((a-b).norm()/b.norm()), float((a-b).abs().max()/b.abs().max()), float((a-b).mean()/b.abs().mean())
# img4 vs im (s2d)
r = im.view(B,57,4,76,4,3).permute(0,1,3,2,4,5).reshape(B,57,76,48)
print("img4", l2(n.img4[...,:48], r))
c = OM.coarse(p64, im, None, False)
print("coarse", l2(n.coarse.view(B,55,74,1), c))
t = T.conv2d(im, p64["fine/first/conv2d/kernel"], p64["fine/first/conv2d/bias"], 2, "valid", True)
t = T.max_pool_2x2(t)
print("fine/first pooled", l2(n.cat[...,:63], t))
cat = torch.cat([t, c], -1)
f2 = T.conv2d(cat, p64["fine/second/conv2d/kernel"], p64["fine/second/conv2d/bias"], 1, "same", True)
print("fine/second", l2(n.f2, f2))
f3 = T.conv2d(f2, p64["fine/third/kernel"], p64["fine/third/bias"], 1, "same", False)
print("fine", l2(n.fine.view(B,55,74,1), f3))
# f2 computed by the GPU from exact inputs?  feed oracle cat into GPU conv
from ann3depth_b200 import ops
ctx = n.ctx
y = ctx.conv2d_fwd(n.d_f2.__class__.from_buffer_copy(n.d_f2), cat.float().cuda().contiguous(), n.w("fine/second/conv2d/kernel"), n.bias("fine/second/conv2d"), relu=True)
print("fine/second from exact input (B rows)", l2(y, f2))
