"""Oracle restatement of the reference MSDN graph (test infrastructure only).

Follows ``/root/reference/src/models.py:203-367`` (class ``_MultiScaleDeepNetwork``).
PARITY UNPINNED by the reference (no tests / golden vectors there); pinned here by the
known-answer tests in ``tests/test_oracle_msdn.py``.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import tf1_ops as T

# (name, shape) in TF variable-creation order.  Names are the TF variable names produced by
# the scopes in src/models.py:208-251 (tf.layers adds '/kernel' and '/bias').
COARSE_CONV = [
    ("coarse/conv/conv2d_0", (11, 11, 3, 96)),     # :211  11x11 s4 VALID
    ("coarse/conv/conv2d_1", (5, 5, 96, 256)),     # :214  5x5 SAME
    ("coarse/conv/conv2d_2", (3, 3, 256, 384)),    # :217
    ("coarse/conv/conv2d_3", (3, 3, 384, 384)),    # :219
    ("coarse/conv/conv2d_4", (3, 3, 384, 256)),    # :222  3x3 s2 VALID
]
COARSE_DENSE = [
    ("coarse/dense/dense_0", (12288, 4096)),       # :228
    ("coarse/dense/dense_1", (4096, 55 * 74)),     # :231
]
FINE = [
    ("fine/first/conv2d", (9, 9, 3, 63)),          # :241  9x9 s2 VALID
    ("fine/second/conv2d", (5, 5, 64, 64)),        # :247  5x5 SAME on concat(63, coarse)
    ("fine/third", (5, 5, 64, 1)),                 # :250  5x5 SAME linear
]
ALL_LAYERS = COARSE_CONV + COARSE_DENSE + FINE

# The four Adam instances of src/models.py:318-345: (name, lr, variable-scope prefixes)
ADAM_GROUPS = OrderedDict([
    ("CoarseConv", (0.001, ("coarse/conv",))),
    ("CoarseDense", (0.1, ("coarse/dense",))),
    ("FineA", (0.001, ("fine/first", "fine/third"))),
    ("FineB", (0.01, ("fine/second",))),
])
ADAM_BETA1 = 0.9     # 'momentum' argument, src/models.py:319-338
ADAM_BETA2 = 1.0     # third positional argument of AdamOptimizer, src/models.py:309
ADAM_EPS = 1e-8      # TF default

IN_H, IN_W = 228, 304        # src/models.py:282
OUT_H, OUT_W = 55, 74        # src/models.py:283
N_PIX = OUT_H * OUT_W        # 4070, the constant 74*55 of src/models.py:269


def param_names():
    names = []
    for n, _ in ALL_LAYERS:
        names += [n + "/kernel", n + "/bias"]
    return names


def init_params(seed=1, dtype=torch.float64, bias_range=0.0):
    """glorot-uniform kernels, zero biases (tf.layers defaults).  ``bias_range`` > 0 draws
    biases from U(-r, r) instead so that the bias paths are exercised by parity tests."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in ALL_LAYERS:
        p[name + "/kernel"] = T.glorot_uniform_(shape, g, dtype)
        if bias_range > 0:
            p[name + "/bias"] = ((torch.rand(shape[-1], generator=g, dtype=torch.float64) * 2 - 1)
                                 * bias_range).to(dtype)
        else:
            p[name + "/bias"] = torch.zeros(shape[-1], dtype=dtype)
    return p


def num_params(p=None):
    if p is None:
        tot = 0
        for _, shape in ALL_LAYERS:
            n = 1
            for s in shape:
                n *= s
            tot += n + shape[-1]
        return tot
    return sum(v.numel() for v in p.values())


def _ident(x):
    return x


def bf16_round(x):
    """Round-trip through bfloat16: emulates the storage precision of the CUDA path."""
    return x.to(torch.bfloat16).to(x.dtype)


def coarse(p, images, dropout_mask=None, train=True, q=_ident):
    """src/models.py:208-236.  images: resized [B,228,304,3].  Returns [B,55,74,1]."""
    def k(n):
        return q(p[n + "/kernel"])

    def b(n):
        return p[n + "/bias"]
    # where a pool follows, q() is applied after it: rounding is monotone so the forward value is the
    # same, and the arg-max routing is decided on unrounded values as in the CUDA path
    t = T.conv2d(images, k("coarse/conv/conv2d_0"), b("coarse/conv/conv2d_0"), 4, "valid", True)
    t = q(T.max_pool_2x2(t))
    t = T.conv2d(t, k("coarse/conv/conv2d_1"), b("coarse/conv/conv2d_1"), 1, "same", True)
    t = q(T.max_pool_2x2(t))
    t = q(T.conv2d(t, k("coarse/conv/conv2d_2"), b("coarse/conv/conv2d_2"), 1, "same", True))
    t = q(T.conv2d(t, k("coarse/conv/conv2d_3"), b("coarse/conv/conv2d_3"), 1, "same", True))
    t = q(T.conv2d(t, k("coarse/conv/conv2d_4"), b("coarse/conv/conv2d_4"), 2, "valid", True))
    t = t.reshape(t.shape[0], -1)                                     # :225 NHWC row-major
    t = q(T.dense(t, k("coarse/dense/dense_0"), b("coarse/dense/dense_0"), "relu"))
    if train:                                                         # :230
        if dropout_mask is None:
            raise ValueError("train=True needs an explicit dropout keep-mask [B,4096]")
        t = q(T.dropout(t, dropout_mask, 0.5))
    t = T.dense(t, k("coarse/dense/dense_1"), b("coarse/dense/dense_1"), None)
    return t.reshape(-1, OUT_H, OUT_W, 1)                             # :234


def fine(p, images, coarse_map, q=_ident):
    """src/models.py:238-253."""
    def k(n):
        return q(p[n + "/kernel"])

    def b(n):
        return p[n + "/bias"]
    t = T.conv2d(images, k("fine/first/conv2d"), b("fine/first/conv2d"), 2, "valid", True)
    t = q(T.max_pool_2x2(t))
    t = torch.cat([t, q(coarse_map)], dim=-1)                         # :246
    t = q(T.conv2d(t, k("fine/second/conv2d"), b("fine/second/conv2d"), 1, "same", True))
    t = T.conv2d(t, k("fine/third"), b("fine/third"), 1, "same", False)
    return t


def silog_loss(outputs, targets):
    """src/models.py:255-275.  NaN logs (argument < 0 after +eps) are replaced by 0.
    Note the l2 term is a SUM over the 4070 pixels (not a mean) and the batch dim is averaged."""
    o = outputs.reshape(outputs.shape[0], -1)
    t = targets.reshape(targets.shape[0], -1)
    eps, lambd = 1e-8, 0.5
    lo = torch.log(o + eps)
    lo = torch.where(torch.isnan(lo), torch.zeros_like(lo), lo)
    lt = torch.log(t + eps)
    lt = torch.where(torch.isnan(lt), torch.zeros_like(lt), lt)
    d = lo - lt
    l2 = (d * d).sum(1)
    si = d.sum(1) ** 2
    return (l2 - lambd / (74 * 55) * si).mean()


def preprocess(images, depths, q=_ident):
    """src/models.py:281-283."""
    return q(T.resize_bilinear_tf1(images, IN_H, IN_W)), T.resize_bilinear_tf1(depths, OUT_H, OUT_W)


def forward(p, images, depths, dropout_mask=None, train=True, q=_ident):
    """src/models.py:277-290.  Returns dict(coarse, fine, loss_coarse, loss_fine, images, depths)."""
    im, dp = preprocess(images, depths, q)
    c = coarse(p, im, dropout_mask, train, q)
    f = fine(p, im, c, q)
    return dict(coarse=c, fine=f, loss_coarse=silog_loss(c, dp), loss_fine=silog_loss(f, dp),
                images=im, depths=dp)


def phase_of(global_step: int, batchsize: int) -> int:
    """src/models.py:301-305,347-364: phase 1 = coarse, 2 = fine, 3 = idle."""
    steps_coarse = 2000000 // batchsize
    steps_fine = 1500000 // batchsize
    if global_step < steps_coarse:
        return 1
    if global_step < steps_coarse + steps_fine:
        return 2
    return 3


def group_of(name: str) -> str:
    for gname, (_, prefixes) in ADAM_GROUPS.items():
        if any(name.startswith(px + "/") for px in prefixes):
            return gname
    raise KeyError(name)


def grads(p, images, depths, dropout_mask, which="coarse", q=_ident, copy=True):
    """Gradients as ``optimizer.compute_gradients(loss, var_list)`` would return them
    (src/models.py:314): 'coarse' -> d loss_coarse / d coarse vars; 'fine' -> d loss_fine /
    d fine vars (coarse map treated as a constant input, since only fine variables are in
    var_list); 'all' -> both, for the all-parameter gradient-parity configuration."""
    # copy=False: the leaves alias the caller's tensors (no 283 MB clone per step; the timed CPU arm of bench.py)
    leaves = OrderedDict((n, (v.detach().clone() if copy else v.detach()).requires_grad_(True)) for n, v in p.items())
    out = forward(leaves, images, depths, dropout_mask, True, q)
    res = OrderedDict()
    if which in ("coarse", "all"):
        names = [n for n in leaves if n.startswith("coarse/")]
        g = torch.autograd.grad(out["loss_coarse"], [leaves[n] for n in names], retain_graph=True)
        res.update(zip(names, g))
    if which in ("fine", "all"):
        names = [n for n in leaves if n.startswith("fine/")]
        g = torch.autograd.grad(out["loss_fine"], [leaves[n] for n in names])
        res.update(zip(names, g))
    fwd = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    return res, fwd


class TrainState:
    """Variables + Adam slots + global_step, as the TF graph of src/models.py:301-367 holds them.
    Each Adam instance has its own beta-power accumulators, advanced only when it is applied."""

    def __init__(self, params, beta2=ADAM_BETA2):
        self.p = OrderedDict((n, v.clone()) for n, v in params.items())
        self.m = OrderedDict((n, torch.zeros_like(v)) for n, v in params.items())
        self.v = OrderedDict((n, torch.zeros_like(v)) for n, v in params.items())
        self.t = {g: 0 for g in ADAM_GROUPS}
        self.global_step = 0
        self.beta2 = beta2


def train_step(state: TrainState, images, depths, dropout_mask, q=_ident, inplace=False):
    """One ``session.run(model_op)`` (src/ann3depth.py:126-127): forward of both stacks, then
    only the taken tf.case branch's gradient/apply ops (src/models.py:353-359).
    inplace=True updates variables and Adam slots in place, as TF does (no per-step copies of the 283 MB state)."""
    B = images.shape[0]
    ph = phase_of(state.global_step, B)
    if ph == 3:
        out = forward(state.p, images, depths, dropout_mask, True, q)
        state.global_step += 1
        return out, ph
    which = "coarse" if ph == 1 else "fine"
    g, out = grads(state.p, images, depths, dropout_mask, which, q, copy=not inplace)
    for gname in (("CoarseConv", "CoarseDense") if ph == 1 else ("FineA", "FineB")):
        state.t[gname] += 1
    for n, gr in g.items():
        gname = group_of(n)
        lr = ADAM_GROUPS[gname][0]
        upd = T.tf_adam_update_ if inplace else T.tf_adam_update
        state.p[n], state.m[n], state.v[n] = upd(
            state.p[n], gr, state.m[n], state.v[n], state.t[gname], lr,
            ADAM_BETA1, state.beta2, ADAM_EPS)
    state.global_step += 1
    return out, ph
