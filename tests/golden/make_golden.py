"""Regenerates tests/golden/*.pt from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference (TensorFlow 1.3) cannot be imported here and ships no vectors of its own, so these are
SELF-GENERATED regression pins of the oracle restatement (see oracle/__init__.py: "parity unpinned").
Inputs are stored alongside the outputs so that neither RNG nor torch version drift matters."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dcnf as OD   # noqa: E402
from oracle import msdn as OM   # noqa: E402
from oracle import tf1_ops as T  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(1234)

# resize (TF1 legacy bilinear), down and up
x = torch.rand(1, 12, 16, 3, generator=g, dtype=torch.float64)
torch.save({"x": x, "down_5x7": T.resize_bilinear_tf1(x, 5, 7), "up_20x33": T.resize_bilinear_tf1(x, 20, 33)},
           os.path.join(OUT, "resize.pt"))

# scale-invariant log loss + gradient (with negative outputs and k/255 targets incl. zeros)
out = (torch.rand(3, 4070, generator=g, dtype=torch.float64) * 1.4 - 0.4).requires_grad_(True)
tar = torch.round(torch.rand(3, 4070, generator=g, dtype=torch.float64) * 255) / 255
loss = OM.silog_loss(out, tar)
(grad,) = torch.autograd.grad(loss, out)
torch.save({"out": out.detach(), "tar": tar, "loss": loss.detach(), "grad": grad}, os.path.join(OUT, "silog_loss.pt"))

# TF-Adam, reference configuration (beta2 = 1) and a sane one
w, gr = torch.randn(257, generator=g, dtype=torch.float64), torch.randn(257, generator=g, dtype=torch.float64) * 1e-2
m, v = torch.randn(257, generator=g, dtype=torch.float64) * 1e-3, torch.rand(257, generator=g, dtype=torch.float64) * 1e-4
ad = {"w": w, "g": gr, "m": m, "v": v}
for b2 in (1.0, 0.999):
    ad[f"out_beta2_{b2}"] = T.tf_adam_update(w, gr, m, v, 3, 0.1, 0.9, b2, 1e-8)
torch.save(ad, os.path.join(OUT, "adam.pt"))

# CRF
z = torch.rand(2, 48, 1, generator=g, dtype=torch.float64)
y = torch.rand(2, 48, 1, generator=g, dtype=torch.float64)
r = torch.rand(2, 48, 1, generator=g, dtype=torch.float64)
A = OD.build_A(r)
torch.save({"z": z, "y": y, "r": r, "ystar": OD.crf_map(A, z), "logdet": torch.linalg.slogdet(A)[1],
            "nll_stable": torch.stack([OD.nll_stable(A[i:i + 1], y[i:i + 1], z[i:i + 1]) for i in range(2)]),
            "nll_naive": torch.stack([OD.nll_naive(A[i:i + 1], y[i:i + 1] * .05, z[i:i + 1] * .05) for i in range(2)])},
           os.path.join(OUT, "crf.pt"))

# MSDN forward at B = 1 (float32 params from seed 1, bias_range .05): outputs only (params are 283 MB)
p = OM.init_params(1, torch.float32, bias_range=0.05)
gi = torch.Generator().manual_seed(0)
im = torch.rand(1, 480, 640, 3, generator=gi)
dp = torch.rand(1, 55, 73, 1, generator=gi) * 0.95 + 0.05
mask = (torch.rand(1, 4096, generator=torch.Generator().manual_seed(2)) < 0.5).float()
f = OM.forward({k: t.double() for k, t in p.items()}, im.double(), dp.double(), mask.double(), True)
torch.save({"coarse": f["coarse"].float(), "fine": f["fine"].float(), "loss_coarse": f["loss_coarse"],
            "loss_fine": f["loss_fine"], "param_checksum": sum(float(t.double().sum()) for t in p.values())},
           os.path.join(OUT, "msdn_forward_b1.pt"))
print("golden vectors written to", OUT)
