"""BASELINE.json config 4: DCNF train step, batch 16 (768 patches), and the CRF-only micro-benchmark."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ann3depth_b200 import models
from ann3depth_b200.init import glorot_params
from ann3depth_b200.dcnf import pair_indices

res = {}
B = 16
g = torch.Generator().manual_seed(3)
images = torch.rand(B, 480, 640, 3, generator=g).cuda()
depths = (torch.rand(B, 480, 640, 1, generator=g) * 0.95 + 0.05).cuda()
op = models.dcnf(images, depths, train=True)
pp = glorot_params(5, "dcnf")
pp["pairwise/pairwise_layers/dense/kernel"].abs_()
op.net.load_params(pp)
for _ in range(2):
    op.run(use_graph=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n):
    op.run(use_graph=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
res["dcnf_train_step_b16"] = {"ms": ms, "images_per_s": B / ms * 1e3, "unary_fwd_tflop": 2.672e9 * 768 / 1e12,
                              "loss": float(op.net.loss), "status_max": int(op.net.status.max())}
print(res["dcnf_train_step_b16"], flush=True)
ctx = op.net.ctx
pl, pr = pair_indices()
pl = torch.tensor(pl, dtype=torch.int32, device="cuda")
pr = torch.tensor(pr, dtype=torch.int32, device="cuda")
for nb in (16, 256, 4096, 65536):
    z, y, r = torch.rand(nb, 48, device="cuda"), torch.rand(nb, 48, device="cuda"), torch.rand(nb, 48, device="cuda")
    ctx.crf(z, y, r, pl, pr)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ctx.crf(z, y, r, pl, pr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[f"crf_b{nb}"] = {"ms": ms, "graphs_per_s": nb / ms * 1e3}
    print(nb, res[f"crf_b{nb}"], flush=True)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "dcnf_bench.json")
json.dump(res, open(out, "w"), indent=1)
