"""Oracle restatement of the reference DCNF graph (test infrastructure only).

Follows ``/root/reference/src/models.py:9-200`` (class ``_DistributedConvolutionalNeuralFields``).
PARITY UNPINNED by the reference; pinned by the known-answer tests in
``tests/test_oracle_dcnf.py`` (pair graph, A = I for r = 0, float64 solves).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import tf1_ops as T

PATCH = (100, 100)      # src/models.py:15
SP = (40, 40)           # src/models.py:16
GAMMA = 1.0             # src/models.py:17
EPS = 1e-7              # src/models.py:18
H, W = 240, 320         # src/models.py:180-181
SGD_LR = 0.1            # src/models.py:198

UNARY_LAYERS = [
    ("unary/unary_layers/conv2d", (11, 11, 3, 64)),      # :64
    ("unary/unary_layers/conv2d_1", (5, 5, 64, 256)),    # :67
    ("unary/unary_layers/conv2d_2", (3, 3, 256, 256)),   # :69
    ("unary/unary_layers/conv2d_3", (3, 3, 256, 256)),   # :71
    ("unary/unary_layers/conv2d_4", (3, 3, 256, 256)),   # :72
    ("unary/unary_layers/dense", (12544, 128)),          # :80
    ("unary/unary_layers/dense_1", (128, 16)),           # :81
    ("unary/unary_layers/dense_2", (16, 1)),             # :82
]
PAIRWISE_LAYERS = [("pairwise/pairwise_layers/dense", (2, 1))]   # :93
ALL_LAYERS = UNARY_LAYERS + PAIRWISE_LAYERS


def num_superpixels(h=H, w=W):
    """src/models.py:32-35."""
    return math.ceil(h / SP[0]), math.ceil(w / SP[1])


def pair_indices(h=H, w=W):
    """src/models.py:20-30: checkerboard of interior centres x 4 neighbours (directed pairs)."""
    max_rows, max_cols = num_superpixels(h, w)
    left, right = [], []
    for row in range(1, max_rows - 1):
        for col in range(2 - (row & 1), max_cols - 1, 2):
            pixel = row * max_cols + col
            for addend in [-max_cols, max_cols, -1, 1]:
                left.append(pixel)
                right.append(pixel + addend)
    return left, right


def init_params(seed=5, dtype=torch.float64, bias_range=0.0, pairwise_nonneg=False):
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in ALL_LAYERS:
        p[name + "/kernel"] = T.glorot_uniform_(shape, g, dtype)
        if bias_range > 0:
            p[name + "/bias"] = ((torch.rand(shape[-1], generator=g, dtype=torch.float64) * 2 - 1)
                                 * bias_range).to(dtype)
        else:
            p[name + "/bias"] = torch.zeros(shape[-1], dtype=dtype)
    if pairwise_nonneg:   # keeps A SPD (the reference has no such constraint, :92-93)
        p["pairwise/pairwise_layers/dense/kernel"].abs_()
    return p


def num_unary_params():
    tot = 0
    for _, shape in UNARY_LAYERS:
        n = 1
        for s in shape:
            n *= s
        tot += n + shape[-1]
    return tot


def superpixels(x):
    """src/models.py:37-48 -> [B, n, 1600, C]."""
    p = T.extract_image_patches_same(x, SP, SP)                  # [B,r,c,40,40,C]
    b = x.shape[0]
    return p.reshape(b, -1, SP[0] * SP[1], x.shape[-1])


def patches(x):
    """src/models.py:50-59 -> [B, n, 100, 100, C]."""
    p = T.extract_image_patches_same(x, PATCH, SP)
    b = x.shape[0]
    return p.reshape(b, -1, PATCH[0], PATCH[1], x.shape[-1])


def unary_part_patch(p, patch_batch, q=lambda t: t):
    """src/models.py:61-83 on [n,100,100,3] -> [n,1]."""
    L = "unary/unary_layers/"

    def k(n):
        return q(p[L + n + "/kernel"])

    def b(n):
        return p[L + n + "/bias"]
    # q() marks the bf16 storage points of the CUDA path (after the pool where one follows: rounding is
    # monotone, and the arg-max routing is decided on unrounded values)
    t = T.conv2d(patch_batch, k("conv2d"), b("conv2d"), 1, "valid", True)
    t = q(T.max_pool_2x2(t))
    t = T.conv2d(t, k("conv2d_1"), b("conv2d_1"), 1, "valid", True)
    t = q(T.max_pool_2x2(t))
    t = q(T.conv2d(t, k("conv2d_2"), b("conv2d_2"), 1, "valid", True))
    t = q(T.conv2d(t, k("conv2d_3"), b("conv2d_3"), 1, "valid", True))
    t = T.conv2d(t, k("conv2d_4"), b("conv2d_4"), 1, "valid", True)
    t = q(T.max_pool_2x2(t))
    t = t.reshape(t.shape[0], -1)
    t = q(T.dense(t, k("dense"), b("dense"), "relu"))
    t = q(T.dense(t, k("dense_1"), b("dense_1"), "sigmoid"))
    t = T.dense(t, k("dense_2"), b("dense_2"), None)
    return t


def unary_part(p, images, q=lambda t: t):
    """src/models.py:85-89 -> z [B, n, 1]."""
    pt = patches(images)
    b, n = pt.shape[:2]
    z = unary_part_patch(p, pt.reshape(b * n, *pt.shape[2:]), q)
    return z.reshape(b, n, 1)


def color_histogram(sp):
    """src/models.py:95-100 for one tile [1600,3] -> [256]."""
    scale = torch.tensor([16777216.0, 65536.0, 256.0], dtype=sp.dtype)
    values = (sp * scale).sum(-1)
    return T.histogram_fixed_width(values, 0.0, 16777216.0, 256)


def similarity(features, pairs):
    """src/models.py:102-106: exp(-gamma * ||f_left - f_right||_2) -> [B, n_pairs]."""
    left = features[:, pairs[0]]
    right = features[:, pairs[1]]
    return torch.exp(-GAMMA * torch.linalg.vector_norm(left - right, dim=2))


def pairwise_features(images):
    """src/models.py:110-125 -> similarities [B, n_pairs, 2] (colour-mean, histogram)."""
    sp = superpixels(images)                                      # [B,n,1600,3]
    pairs = pair_indices(images.shape[1], images.shape[2])
    hist = torch.stack([torch.stack([color_histogram(t) for t in bt]) for bt in sp])
    cdiff = similarity(sp.mean(-1), pairs)
    hdiff = similarity(hist, pairs)
    return torch.stack([cdiff, hdiff], dim=-1)


def pairwise_part(p, images):
    """src/models.py:108-127 -> r [B, n_pairs, 1]."""
    sims = pairwise_features(images)
    n = "pairwise/pairwise_layers/dense"
    return T.dense(sims, p[n + "/kernel"], p[n + "/bias"], None)


def build_A(r, h=H, w=W):
    """src/models.py:136-152: R[l,r] = R[r,l] = r_k ; A = I + diag(R 1) - R.  r: [B,n_pairs,1].
    The reference sizes R by the number of pairs (:149), valid only when #pairs == #nodes."""
    rows, cols = num_superpixels(h, w)
    n = rows * cols
    left, right = pair_indices(h, w)
    b = r.shape[0]
    R = torch.zeros(b, n, n, dtype=r.dtype)
    rv = r.reshape(b, -1)
    R[:, left, right] = rv
    R[:, right, left] = rv
    D = torch.diag_embed(R.sum(2))
    return torch.eye(n, dtype=r.dtype) + D - R


def tile_means(depths):
    """src/models.py:131-132 -> y [B, n, 1]."""
    return superpixels(depths).mean(2)


def crf_terms(A, y, z):
    """Energy, log det A and z^T A^-1 z used by both the naive and the stable NLL."""
    yT, zT = y.transpose(1, 2), z.transpose(1, 2)
    energy = (yT @ A @ y - 2 * zT @ y + zT @ z).reshape(-1)       # :159
    return energy


def nll_naive(A, y, z):
    """src/models.py:157-174 evaluated literally (exp/det/inverse; saturates in float32)."""
    n = A.shape[-1]
    zT = z.transpose(1, 2)
    energy = crf_terms(A, y, z)
    fac = math.pi ** (n / 2) / (torch.linalg.det(A) ** 0.5 + EPS)           # :163-164
    invA = torch.linalg.inv(A) + EPS                                        # :165 (elementwise +eps)
    ex = torch.exp((zT @ invA @ z - zT @ z).reshape(-1))                    # :166
    Z = fac * ex + EPS                                                      # :167
    loss = -torch.log(torch.exp(-energy) / Z + EPS)                         # :171
    return loss.mean()                                                      # :174


def nll_stable(A, y, z):
    """The same NLL in closed form without the eps terms:
    E + (n/2) log pi - 0.5 log det A + z^T A^-1 z - z^T z  (Liu et al. 2015, eq. 9-11)."""
    n = A.shape[-1]
    zT = z.transpose(1, 2)
    energy = crf_terms(A, y, z)
    logdet = torch.linalg.slogdet(A)[1]
    sol = torch.linalg.solve(A, z)
    quad = (zT @ sol).reshape(-1)
    zz = (zT @ z).reshape(-1)
    return (energy + 0.5 * n * math.log(math.pi) - 0.5 * logdet + quad - zz).mean()


def crf_map(A, z):
    """MAP prediction y* = A^-1 z (Liu et al. eq. 12; north-star 'y = A^-1 z')."""
    return torch.linalg.solve(A, z)


def forward(p, images, depths, q=lambda t: t, stable=False, grad_through_r=False):
    """src/models.py:179-191.  Returns dict(z, r, A, y, loss, output, images, depths)."""
    im = T.resize_bilinear_tf1(images, H, W)
    dp = T.resize_bilinear_tf1(depths, H, W)
    z = unary_part(p, q(im), q)
    r = pairwise_part(p, im)
    # ASSUMPTION: tf.scatter_nd_update has no registered gradient in TF 1.3, so no gradient
    # reaches r / pairwise_layers in the reference (src/models.py:138-141).
    A = build_A(r if grad_through_r else r.detach())
    y = tile_means(dp)
    loss = (nll_stable if stable else nll_naive)(A, y, z)
    rows, cols = num_superpixels()
    output = T.resize_bilinear_tf1(z.reshape(-1, rows, cols, 1), H, W)      # :187-191
    return dict(z=z, r=r, A=A, y=y, loss=loss, output=output, images=im, depths=dp)


def grads(p, images, depths, q=lambda t: t, stable=True):
    leaves = OrderedDict((n, v.detach().clone().requires_grad_(True)) for n, v in p.items())
    out = forward(leaves, images, depths, q, stable)
    names = [n for n in leaves if n.startswith("unary/")]
    g = torch.autograd.grad(out["loss"], [leaves[n] for n in names])
    return OrderedDict(zip(names, g)), {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
