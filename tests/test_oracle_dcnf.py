"""Known-answer tests pinning the DCNF oracle (KA2, KA6, KA7, KA10)."""
import math

import torch

from oracle import dcnf as D


def test_ka2_unary_param_count():
    assert D.num_unary_params() == 3811233


def test_ka6_pair_graph():
    left, right = D.pair_indices()
    assert len(left) == len(right) == 48
    assert sorted(set(left)) == [9, 11, 13, 18, 20, 22, 25, 27, 29, 34, 36, 38]
    touched = set(left) | set(right)
    assert sorted(set(range(48)) - touched) == [0, 2, 4, 6, 7, 15, 16, 31, 32, 40, 41, 43, 45, 47]
    assert len({frozenset(e) for e in zip(left, right)}) == 48
    assert D.num_superpixels() == (6, 8)


def test_ka7_crf_identity_and_solve():
    b, n = 3, 48
    g = torch.Generator().manual_seed(7)
    z = torch.rand(b, n, 1, generator=g, dtype=torch.float64)
    y = torch.rand(b, n, 1, generator=g, dtype=torch.float64)
    A0 = D.build_A(torch.zeros(b, 48, 1, dtype=torch.float64))
    assert torch.equal(A0, torch.eye(n, dtype=torch.float64).expand(b, n, n))
    assert torch.allclose(D.crf_map(A0, z), z)
    e = D.crf_terms(A0, y, z)
    assert torch.allclose(e, ((y - z) ** 2).sum((1, 2)))
    r = torch.rand(b, 48, 1, generator=g, dtype=torch.float64)
    A = D.build_A(r)
    assert torch.allclose(A, A.transpose(1, 2))
    assert torch.allclose(A.sum(2), torch.ones(b, n, dtype=torch.float64))   # rows of D - R sum to 0
    assert float(torch.linalg.eigvalsh(A).min()) >= 1.0 - 1e-9                 # SPD for r >= 0
    ystar = D.crf_map(A, z)
    assert torch.allclose(A @ ystar, z, atol=1e-12)
    # the naive form is the stable closed form pushed through exp / +eps / -log (it saturates at
    # -log(1e-7) = 16.118 whenever exp(-nll) << 1e-7): check that relation per sample
    for i in range(b):
        st = float(D.nll_stable(A[i:i + 1], y[i:i + 1] * 0.05, z[i:i + 1] * 0.05))
        nv = float(D.nll_naive(A[i:i + 1], y[i:i + 1] * 0.05, z[i:i + 1] * 0.05))
        assert abs(nv - (-math.log(math.exp(-st) + D.EPS))) < 1e-5
    assert abs(float(D.nll_naive(A, y * 50, z * 50)) - 16.118095) < 1e-4


def test_ka10_histogram():
    red = torch.zeros(1600, 3, dtype=torch.float64)
    red[:, 0] = 1.0
    h = D.color_histogram(red)
    assert float(h[255]) == 1600 and float(h.sum()) == 1600
    black = torch.zeros(1600, 3, dtype=torch.float64)
    h = D.color_histogram(black)
    assert float(h[0]) == 1600


def test_patches_and_tiles_shapes():
    x = torch.rand(2, 240, 320, 3, dtype=torch.float64)
    assert D.patches(x).shape == (2, 48, 100, 100, 3)
    sp = D.superpixels(x)
    assert sp.shape == (2, 48, 1600, 3)
    # tile 9 = row 1, col 1 -> pixels [40:80, 40:80]
    assert torch.equal(sp[0, 9].reshape(40, 40, 3), x[0, 40:80, 40:80])
    # patch 0 is centred on tile 0 with 30 px of zero padding on top/left
    pt = D.patches(x)
    assert float(pt[0, 0, :30].abs().max()) == 0 and torch.equal(pt[0, 0, 30:, 30:], x[0, :70, :70])
