"""Regression pins: the oracle against the committed vectors (CPU), and liba3d kernels against the same
vectors (GPU).  The vectors are self-generated (tests/golden/make_golden.py): the reference has none."""
import os

import pytest
import torch

from oracle import dcnf as OD
from oracle import msdn as OM
from oracle import tf1_ops as T

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def test_oracle_matches_golden_small_ops():
    d = load("resize.pt")
    assert torch.equal(T.resize_bilinear_tf1(d["x"], 5, 7), d["down_5x7"])
    assert torch.equal(T.resize_bilinear_tf1(d["x"], 20, 33), d["up_20x33"])
    d = load("silog_loss.pt")
    o = d["out"].clone().requires_grad_(True)
    loss = OM.silog_loss(o, d["tar"])
    assert torch.allclose(loss, d["loss"], rtol=1e-12)
    assert torch.allclose(torch.autograd.grad(loss, o)[0], d["grad"], rtol=1e-10, atol=1e-14)
    d = load("adam.pt")
    for b2 in (1.0, 0.999):
        got = T.tf_adam_update(d["w"], d["g"], d["m"], d["v"], 3, 0.1, 0.9, b2, 1e-8)
        for a, b in zip(got, d[f"out_beta2_{b2}"]):
            assert torch.allclose(a, b, rtol=1e-12, atol=1e-15)
    d = load("crf.pt")
    A = OD.build_A(d["r"])
    assert torch.allclose(OD.crf_map(A, d["z"]), d["ystar"], atol=1e-12)
    assert torch.allclose(torch.linalg.slogdet(A)[1], d["logdet"], atol=1e-12)


def test_oracle_matches_golden_msdn_forward():
    d = load("msdn_forward_b1.pt")
    p = OM.init_params(1, torch.float32, bias_range=0.05)
    assert abs(sum(float(t.double().sum()) for t in p.values()) - d["param_checksum"]) < 1e-6
    gi = torch.Generator().manual_seed(0)
    im = torch.rand(1, 480, 640, 3, generator=gi)
    dp = torch.rand(1, 55, 73, 1, generator=gi) * 0.95 + 0.05
    mask = (torch.rand(1, 4096, generator=torch.Generator().manual_seed(2)) < 0.5).float()
    f = OM.forward({k: t.double() for k, t in p.items()}, im.double(), dp.double(), mask.double(), True)
    assert torch.allclose(f["coarse"].float(), d["coarse"], atol=1e-5)
    assert torch.allclose(f["fine"].float(), d["fine"], atol=1e-5)
    assert abs(float(f["loss_coarse"]) - float(d["loss_coarse"])) < 1e-6 * abs(float(d["loss_coarse"]))


@pytest.mark.gpu
def test_kernels_match_golden():
    from ann3depth_b200 import ops
    c = ops.Context(0)
    dev = "cuda:0"
    d = load("resize.pt")
    y = c.resize_bilinear_tf1(d["x"].float().to(dev), 5, 7)
    assert float((y.cpu().double() - d["down_5x7"]).abs().max()) < 1e-6
    y = c.resize_bilinear_tf1(d["x"].float().to(dev), 20, 33)
    assert float((y.cpu().double() - d["up_20x33"]).abs().max()) < 1e-6
    d = load("silog_loss.pt")
    loss, _, dout, _ = c.silog_loss(d["out"].float().to(dev), d["tar"].float().to(dev))
    assert abs(float(loss) - float(d["loss"])) / float(d["loss"]) < 1e-4
    assert float((dout.cpu().double() - d["grad"]).abs().max() / d["grad"].abs().max()) < 1e-4
    d = load("adam.pt")
    for b2 in (1.0, 0.999):
        w, g, m, v = (d[k].float().to(dev).clone() for k in ("w", "g", "m", "v"))
        c.adam_tf(w, g, m, v, None, 0.1, 0.9, b2, 1e-8, 3)
        rw, rm, rv = d[f"out_beta2_{b2}"]
        assert float((w.cpu().double() - rw).abs().max()) < 1e-5
        assert float((m.cpu().double() - rm).abs().max()) < 1e-8
    d = load("crf.pt")
    pl, pr = OD.pair_indices()
    res = c.crf(d["z"].float().reshape(2, 48).to(dev), d["y"].float().reshape(2, 48).to(dev),
                d["r"].float().reshape(2, 48).to(dev), torch.tensor(pl, dtype=torch.int32, device=dev),
                torch.tensor(pr, dtype=torch.int32, device=dev))
    assert float((res["ystar"].cpu().double() - d["ystar"].reshape(2, 48)).abs().max()) < 1e-5
    assert float((res["nll"].cpu().double() - d["nll_stable"]).abs().max()) < 1e-3
    torch.cuda.synchronize()
    c.close()
