import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ann3depth_b200 import models
from ann3depth_b200.init import glorot_params

B = 2
g = torch.Generator().manual_seed(0)
images = torch.rand(B, 480, 640, 3, generator=g).cuda()
depths = (torch.rand(B, 55, 73, 1, generator=g) * 0.95 + 0.05).cuda()
mask = (torch.rand(B, 4096, generator=torch.Generator().manual_seed(2)) < 0.5).cuda()
p = glorot_params(1)
p["coarse/dense/dense_1/bias"] += 1.0
res = {}
MODES = {"seq": False, "ovl": True, "ovl2": True, "fine": "fine", "wgrad": "wgrad"}
for mode in MODES:
    op = models.msdn(images, depths, train=True, overlap=MODES[mode])
    op.net.load_params(p)
    op.net.set_dropout_mask(mask)
    op.run(use_graph=False)
    torch.cuda.synchronize()
    res[mode] = (op.net.arena.g.clone(), op.net.arena.m.clone(), op.net)
    for nm in ("g_coarse", "g_d0a", "g_d0", "g_c4a", "g_c4", "g_c3", "d0", "c4", "coarse", "fine"):
        res[mode] += (getattr(op.net, nm).clone(),)
net = res["seq"][2]
for name, s in net.arena.specs.items():
    for k, label in ((0, "g"), (1, "m")):
        a = res["seq"][k][s.offset:s.offset + s.numel]
        b = res["ovl"][k][s.offset:s.offset + s.numel]
        c = res["ovl2"][k][s.offset:s.offset + s.numel]
        d1 = float((a - b).abs().max()); d2 = float((b - c).abs().max()); mx = float(a.abs().max())
        if d1 > 1e-5 * mx or d2 > 1e-5 * mx:
            print(f"{name:34s} {label} max|seq|={mx:.3e} |seq-ovl|={d1:.3e} |ovl-ovl2|={d2:.3e}")
names = ("g_coarse", "g_d0a", "g_d0", "g_c4a", "g_c4", "g_c3", "d0", "c4", "coarse", "fine")
for mode in ("ovl", "fine", "wgrad"):
    print("==", mode)
    for i, nm in enumerate(names):
        a, b = res["seq"][3 + i].float(), res[mode][3 + i].float()
        print(f"   {nm:10s} max|diff|={float((a - b).abs().max()):.3e} (max {float(a.abs().max()):.3e})")
    for name in ("coarse/dense/dense_1/kernel", "coarse/dense/dense_0/kernel", "coarse/dense/dense_0/bias", "coarse/conv/conv2d_4/kernel"):
        s = net.arena.specs[name]
        a = res["seq"][0][s.offset:s.offset + s.numel]; b = res[mode][0][s.offset:s.offset + s.numel]
        print(f"   g[{name}] max|diff|={float((a - b).abs().max()):.3e} (max {float(a.abs().max()):.3e})")
print("done")
