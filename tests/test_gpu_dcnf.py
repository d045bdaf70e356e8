"""End-to-end GPU parity of the DCNF step against the CPU oracle (oracle/dcnf.py) through
`ann3depth_b200.models.dcnf`.  CRF solve tolerance 1e-5 (north star); unary outputs 1e-2 (BF16);
unary gradient cosine >= 0.999 against the oracle evaluated at the BF16 storage points."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dcnf as OD

if torch.cuda.is_available():
    from ann3depth_b200 import models

DEV = "cuda:0"


def make(B=1, seed=3):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(B, 480, 640, 3, generator=g)
    depths = torch.rand(B, 480, 640, 1, generator=g) * 0.95 + 0.05
    p = OD.init_params(5, torch.float32, bias_range=0.05, pairwise_nonneg=True)
    return images, depths, p


def cos(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def test_dcnf_forward_and_gradients():
    B = 1
    images, depths, p = make(B)
    op = models.dcnf(images.to(DEV), depths.to(DEV), train=True, naive_loss=False)
    net = op.net
    net.load_params(p)
    net.forward()
    net.backward()
    torch.cuda.synchronize()
    p64 = {k: v.double() for k, v in p.items()}
    gref, ref = OD.grads(p64, images.double(), depths.double(), q=OD_bf16, stable=True)
    z, zr = net.z.view(B, 48).cpu().double(), ref["z"].reshape(B, 48)
    print("z rel", float((z - zr).abs().max() / zr.abs().max()))
    assert float((z - zr).abs().max() / zr.abs().max()) < 1e-2
    assert float((net.r.cpu().double() - ref["r"].reshape(B, 48)).abs().max()) < 1e-5
    assert float((net.y.cpu().double() - ref["y"].reshape(B, 48)).abs().max()) < 1e-6
    assert int(net.status.abs().max()) == 0
    # CRF solve on the GPU's own z: y* = A^-1 z within 1e-5 of a float64 solve
    A = OD.build_A(net.r.cpu().double().reshape(B, 48, 1))
    ystar = OD.crf_map(A, z.reshape(B, 48, 1)).reshape(B, 48)
    assert float((net.ystar.cpu().double() - ystar).abs().max()) < 1e-5
    lref = float(OD.nll_stable(A, ref["y"], z.reshape(B, 48, 1)))
    assert abs(float(net.loss) - lref) < 1e-3 * max(1.0, abs(lref))
    print("loss", float(net.loss), "oracle (same z)", lref, "oracle e2e", float(ref["loss"]))
    assert abs(float(net.loss) - float(ref["loss"])) < 2e-2 * max(1.0, abs(float(ref["loss"])))
    out = OD.T.resize_bilinear_tf1(z.reshape(B, 6, 8, 1), 240, 320)
    assert float((net.output.cpu().double() - out).abs().max()) < 1e-5
    got = net.export_grads()
    bad = []
    for name, g in gref.items():
        c = cos(got[name], g)
        nr = float(got[name].double().norm() / (g.norm() + 1e-300))
        print(f"{name:36s} cos={c:.6f} norm ratio={nr:.4f}")
        if not (c >= 0.999 and 0.97 < nr < 1.03):
            bad.append(name)
    assert not bad, bad
    assert float(got["pairwise/pairwise_layers/dense/kernel"].abs().max()) == 0.0   # no gradient reaches r


def OD_bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


def test_dcnf_train_step_sgd_and_naive_loss():
    B = 1
    images, depths, p = make(B, seed=4)
    op = models.dcnf(images.to(DEV), depths.to(DEV), train=True)       # reference loss form (naive)
    net = op.net
    net.load_params(p)
    w0 = net.arena.w.clone()
    op.run()
    torch.cuda.synchronize()
    assert op.global_step == 1
    lo, hi = net.arena.group_range("SGD")
    assert torch.allclose(net.arena.w[lo:hi], w0[lo:hi] - 0.1 * net.arena.g[lo:hi], atol=1e-7)
    lo, hi = net.arena.group_range("Pairwise")
    assert torch.equal(net.arena.w[lo:hi], w0[lo:hi])
    # naive loss == -log(exp(-stable) + eps) up to the reference's other eps terms; saturates at 16.118
    A = OD.build_A(net.r.cpu().double().reshape(B, 48, 1))
    zz = net.z.view(B, 48, 1).cpu().double()
    yy = net.y.view(B, 48, 1).cpu().double()
    naive = float(OD.nll_naive(A, yy, zz))
    print("naive loss", float(net.loss), "oracle", naive)
    assert abs(float(net.loss) - naive) < 1e-3 * max(1.0, abs(naive))


def test_dcnf_beyond_reference_options():
    """SURVEY.md 8f N4 (none of this exists in the reference, src/models.py:129-191): a grid with #pairs != #nodes, the
    r >= 0 constraint, the MAP estimate y* = A^-1 z as the prediction, and the gradient INTO the pairwise layer -- against
    float64 linear algebra / autograd on the same z, y and similarities."""
    from ann3depth_b200.dcnf import grid4_pairs
    B = 2
    images, depths, p = make(B, seed=6)
    # tile-wise nearly constant colours: neighbouring tiles differ by ~0.02 per channel, so the colour similarity
    # exp(-||mean_l - mean_r||_2) over the 1600 tile pixels lands in (0.2, 0.9) instead of underflowing to 0
    g = torch.Generator().manual_seed(16)
    tiles = 0.5 + 0.03 * torch.rand(B, 6, 8, 3, generator=g)
    images = tiles.repeat_interleave(80, 1).repeat_interleave(80, 2).contiguous()
    P = "pairwise/pairwise_layers/dense"
    p[P + "/kernel"] = torch.tensor([[0.8], [-0.9]])            # mixed signs ...
    p[P + "/bias"] = torch.tensor([-0.35])                      # ... and an offset: the clamp at 0 is exercised
    op = models.dcnf(images.to(DEV), depths.to(DEV), train=True, naive_loss=False, graph="grid4", r_nonneg=True,
                     train_pairwise=True, predict="map")
    net = op.net
    net.load_params(p)
    assert net.n_pairs == 82 and net.n == 48
    net.forward()
    net.backward()
    torch.cuda.synchronize()
    sims = net.sims.cpu().double()                               # [B,82,2]
    z = net.z.view(B, 48, 1).cpu().double()
    y = net.y.view(B, 48, 1).cpu().double()
    w = p[P + "/kernel"].double().reshape(2).requires_grad_(True)
    b = p[P + "/bias"].double().requires_grad_(True)
    pl, pr = grid4_pairs()

    def build(rv):
        R = torch.zeros(B, 48, 48, dtype=torch.float64)
        R = R.index_put((torch.arange(B)[:, None], torch.tensor(pl)[None], torch.tensor(pr)[None]), rv)
        R = R.index_put((torch.arange(B)[:, None], torch.tensor(pr)[None], torch.tensor(pl)[None]), rv)
        return torch.eye(48, dtype=torch.float64) + torch.diag_embed(R.sum(2)) - R
    r_ref = torch.relu(sims @ w + b)
    assert float((net.r.cpu().double() - r_ref.detach()).abs().max()) < 1e-6
    print("r range", float(r_ref.min()), float(r_ref.max()), "clamped fraction", float((r_ref == 0).double().mean()))
    assert float(r_ref.min()) == 0.0 and float(r_ref.max()) > 0.0        # both branches of the clamp occur
    A = build(r_ref)
    loss = OD.nll_stable(A, y, z)
    gw, gb = torch.autograd.grad(loss, [w, b])
    assert int(net.status.abs().max()) == 0
    ystar = torch.linalg.solve(A.detach(), z).reshape(B, 48)
    assert float((net.ystar.cpu().double() - ystar).abs().max()) < 1e-5          # north star: CRF solve 1e-5
    assert abs(float(net.loss) - float(loss)) < 1e-4 * max(1.0, abs(float(loss)))
    out = OD.T.resize_bilinear_tf1(ystar.reshape(B, 6, 8, 1), 240, 320)
    assert float((net.output.cpu().double() - out).abs().max()) < 1e-4           # the prediction is the upsampled MAP
    got = net.export_grads()
    print("pairwise grads", got[P + "/kernel"].reshape(-1).tolist(), gw.tolist(), got[P + "/bias"].tolist(), gb.tolist())
    assert float((got[P + "/kernel"].double().reshape(2) - gw).abs().max()) < 1e-3 * max(1e-3, float(gw.abs().max()))
    assert abs(float(got[P + "/bias"]) - float(gb)) < 1e-3 * max(1e-3, abs(float(gb)))
    w0 = net.arena.w.clone()
    op.run()
    torch.cuda.synchronize()
    lo, hi = net.arena.group_range("Pairwise")
    assert torch.allclose(net.arena.w[lo:hi], w0[lo:hi] - 0.1 * net.arena.g[lo:hi], atol=1e-7)
    assert not torch.equal(net.arena.w[lo:hi], w0[lo:hi])


def test_dcnf_fully_convolutional_equals_patchwise():
    """The default evaluation runs the unary CNN ONCE per zero-padded image and gathers 7x7 windows for the dense layers;
    unary="patches" runs the reference's literal formulation (48 overlapping 100x100 patches, src/models.py:50-83) on the
    same kernels.  Same forward values (each output is the same dot product), same gradients up to the bf16 rounding of
    activation gradients that the fully convolutional form sums over overlapping patches before rounding."""
    B = 2
    images, depths, p = make(B, seed=8)
    res = {}
    for mode in ("fullconv", "patches"):
        op = models.dcnf(images.to(DEV), depths.to(DEV), train=True, naive_loss=False, unary=mode)
        net = op.net
        net.load_params(p)
        net.forward()
        net.backward()
        torch.cuda.synchronize()
        res[mode] = (net.z.clone(), net.h0.clone(), float(net.loss), net.export_grads())
    zf, zp = res["fullconv"][0], res["patches"][0]
    print("z max diff", float((zf - zp).abs().max()), "h0 max diff", float((res["fullconv"][1].float() - res["patches"][1].float()).abs().max()))
    assert float((zf - zp).abs().max()) <= 1e-3 * float(zp.abs().max())
    assert abs(res["fullconv"][2] - res["patches"][2]) < 1e-3 * max(1.0, abs(res["patches"][2]))
    for name, g in res["patches"][3].items():
        if name.startswith("pairwise"):
            continue
        c = cos(res["fullconv"][3][name], g)
        nr = float(res["fullconv"][3][name].double().norm() / (g.double().norm() + 1e-300))
        print(f"{name:36s} cos={c:.7f} norm ratio={nr:.5f}")
        assert c >= 0.9999 and 0.99 < nr < 1.01, name


def test_dcnf_inference_equals_training_forward():
    """train=False pools the bf16 conv outputs (no routing record): bit-identical predictions (rounding is monotone)."""
    B = 2
    images, depths, p = make(B, seed=9)
    opt = models.dcnf(images.to(DEV), depths.to(DEV), train=True, naive_loss=False)
    opt.net.load_params(p)
    opt.net.forward()
    opi = models.dcnf(images.to(DEV), depths.to(DEV), train=False, naive_loss=False)
    opi.net.load_params(p)
    out = opi.run()
    torch.cuda.synchronize()
    assert opi.net.c1.dtype == torch.bfloat16 and opt.net.c1.dtype == torch.float32
    assert torch.equal(opi.net.z, opt.net.z) and torch.equal(out, opt.net.output)


def test_dcnf_graph_step_equals_eager_step():
    B = 1
    images, depths, p = make(B, seed=10)
    res = {}
    for use_graph in (False, True):
        op = models.dcnf(images.to(DEV), depths.to(DEV), train=True, naive_loss=False)
        op.net.load_params(p)
        for _ in range(3):
            op.run(use_graph=use_graph)
        torch.cuda.synchronize()
        assert op.global_step == 3
        res[use_graph] = (op.net.arena.w.clone(), float(op.net.loss))
    # split-K atomics make two runs differ in the last bits; three SGD steps stay within 1e-5 of the weights' scale
    d = float((res[True][0] - res[False][0]).abs().max())
    print("graph vs eager max weight diff", d, "loss", res[True][1], res[False][1])
    assert d < 1e-4 * float(res[False][0].abs().max())
    assert abs(res[True][1] - res[False][1]) < 1e-3 * max(1.0, abs(res[False][1]))
