"""Parameter layouts: conversion between the reference's TensorFlow variable layouts and the packed
layouts liba3d computes on, and the flat parameter arena (weights / grads / Adam slots / bf16 mirror).

TF side (src/models.py): conv kernels HWIO `[kh, kw, in, out]`, dense kernels `[in, out]`,
variable names `<scope>/kernel`, `<scope>/bias`.
Packed side: conv kernels OHWI `[out][kh][kw'][in']`, dense kernels `[out][in]`; kw'/in'/out may be
zero-padded so the first (3-channel) layers map onto 16-byte pixels groups.
The arena orders variables in BACKWARD order inside each optimizer group, so gradient buckets
become ready front-to-back during the backward pass and each Adam group is one contiguous range.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import torch


@dataclass
class VarSpec:
    name: str            # TF variable name
    tf_shape: tuple
    packed_shape: tuple
    kind: str            # 'conv_kernel' | 'conv_kernel_s2d' | 'dense_kernel' | 'bias'
    group: str           # optimizer group
    offset: int = 0      # element offset in the arena
    size: int = 0        # padded element count (multiple of 8)

    @property
    def numel(self):
        n = 1
        for s in self.packed_shape:
            n *= s
        return n


def pack(spec: VarSpec, t: torch.Tensor) -> torch.Tensor:
    """TF layout -> packed layout (zero padding where the packed shape is larger)."""
    if spec.kind == "conv_kernel":
        kh, kw, ci, co = spec.tf_shape
        K, R, S, Cp = spec.packed_shape
        out = torch.zeros(spec.packed_shape, dtype=t.dtype)
        out[:co, :kh, :kw, :ci] = t.permute(3, 0, 1, 2)
        return out
    if spec.kind == "conv_kernel_s2d":
        # stride-2 filter [kh,kw,ci,co] as the stride-1 filter over the space-to-depth input:
        # packed[co][r'][s'][(di*2+dj)*4 + c] = t[2r'+di][2s'+dj][c][co]  (zero beyond kh, kw, ci)
        kh, kw, ci, co = spec.tf_shape
        K, R2, S2, C16 = spec.packed_shape
        full = torch.zeros(2 * R2, 2 * S2, 4, K, dtype=t.dtype)
        full[:kh, :kw, :ci, :co] = t
        full = full.view(R2, 2, S2, 2, 4, K).permute(5, 0, 2, 1, 3, 4)       # K, r', s', di, dj, c
        return full.reshape(K, R2, S2, 16).contiguous()
    if spec.kind == "conv_kernel_s2d4":
        return pack_s2d(t, 4, spec.packed_shape)
    if spec.kind == "dense_kernel":
        n_in, n_out = spec.tf_shape
        if tuple(spec.packed_shape) == (n_out, n_in):
            return t.t().contiguous()
        out = torch.zeros(spec.packed_shape, dtype=t.dtype)           # output rows padded (dense_1: 4070 -> 4096)
        out[:n_out, :n_in] = t.t()
        return out
    out = torch.zeros(spec.packed_shape, dtype=t.dtype)
    out[:t.numel()] = t
    return out


def pack_s2d(t: torch.Tensor, s: int, packed_shape) -> torch.Tensor:
    """Stride-s filter [kh,kw,ci,co] as the stride-1 filter over the space-to-depth(s) input:
    packed[co][R][S][(dy*s+dx)*ci + c] = t[s*R+dy][s*S+dx][c][co]; zero beyond kh, kw, co and in the channel padding."""
    kh, kw, ci, co = t.shape
    K, R2, S2, Cs = packed_shape
    full = torch.zeros(s * R2, s * S2, ci, K, dtype=t.dtype)
    full[:kh, :kw, :, :co] = t
    full = full.view(R2, s, S2, s, ci, K).permute(5, 0, 2, 1, 3, 4).reshape(K, R2, S2, s * s * ci)
    out = torch.zeros(packed_shape, dtype=t.dtype)
    out[..., :s * s * ci] = full
    return out


def unpack_s2d(p: torch.Tensor, s: int, tf_shape) -> torch.Tensor:
    kh, kw, ci, co = tf_shape
    K, R2, S2, Cs = p.shape
    full = p[..., :s * s * ci].reshape(K, R2, S2, s, s, ci).permute(1, 3, 2, 4, 5, 0).reshape(s * R2, s * S2, ci, K)
    return full[:kh, :kw, :, :co].contiguous()


def embed_pool4(t: torch.Tensor, conv_stride: int = 2, kc: int = 64) -> torch.Tensor:
    """conv(stride) + 2x2/2 max-pool as one conv of stride 2*stride: the four conv outputs of a pool window share
    the receptive field of size k + stride, so filter (a, b) of the window is the original filter shifted by
    (a*stride, b*stride) inside it.  [kh,kw,ci,co] -> [kh+stride, kw+stride, ci, 4*kc], filter index (2a+b)*kc + c."""
    kh, kw, ci, co = t.shape
    big = torch.zeros(kh + conv_stride, kw + conv_stride, ci, 4 * kc, dtype=t.dtype)
    for a in range(2):
        for b in range(2):
            g = 2 * a + b
            big[a * conv_stride:a * conv_stride + kh, b * conv_stride:b * conv_stride + kw, :, g * kc:g * kc + co] = t
    return big


def embed_index_map(spec: "VarSpec", derive) -> torch.Tensor:
    """int32 [G, spec.numel]: position of every canonical (packed) element of `spec` in each copy of the derived
    tensor produced by `derive(tf_tensor) -> list of G flat tensors` (one per copy, zeros elsewhere); -1 = none."""
    n_tf = 1
    for d in spec.tf_shape:
        n_tf *= d
    ids = torch.arange(1, n_tf + 1, dtype=torch.float64).reshape(spec.tf_shape)
    canon = pack(spec, ids).reshape(-1).long()
    canon_pos = torch.full((n_tf + 1,), -1, dtype=torch.long)
    nz = canon.nonzero().reshape(-1)
    canon_pos[canon[nz]] = nz
    copies = derive(ids)
    out = torch.full((len(copies), canon.numel()), -1, dtype=torch.int32)
    for g, big in enumerate(copies):
        big = big.reshape(-1).long()
        pos = big.nonzero().reshape(-1)
        out[g, canon_pos[big[pos]]] = pos.to(torch.int32)
    return out


def unpack(spec: VarSpec, p: torch.Tensor) -> torch.Tensor:
    """packed layout -> TF layout."""
    p = p.reshape(spec.packed_shape)
    if spec.kind == "conv_kernel":
        kh, kw, ci, co = spec.tf_shape
        return p[:co, :kh, :kw, :ci].permute(1, 2, 3, 0).contiguous()
    if spec.kind == "conv_kernel_s2d":
        kh, kw, ci, co = spec.tf_shape
        K, R2, S2, C16 = spec.packed_shape
        full = p.view(K, R2, S2, 2, 2, 4).permute(1, 3, 2, 4, 5, 0).reshape(2 * R2, 2 * S2, 4, K)
        return full[:kh, :kw, :ci, :co].contiguous()
    if spec.kind == "conv_kernel_s2d4":
        return unpack_s2d(p, 4, spec.tf_shape)
    if spec.kind == "dense_kernel":
        n_in, n_out = spec.tf_shape
        return p[:n_out, :n_in].t().contiguous()
    return p[:spec.tf_shape[0]].contiguous()


def keep_mask(spec: VarSpec) -> torch.Tensor | None:
    """uint8 mask of the real (non-padding) entries of a packed variable, or None if unpadded."""
    ones = torch.ones(spec.tf_shape, dtype=torch.uint8)
    m = pack(spec, ones)
    return None if bool(m.all()) else m


def _conv(name, tf_shape, packed, group):
    return [VarSpec(name + "/kernel", tf_shape, packed, "conv_kernel", group),
            VarSpec(name + "/bias", (tf_shape[-1],), (packed[0],), "bias", group)]


def _dense(name, n_in, n_out, group, rows=None):
    return [VarSpec(name + "/kernel", (n_in, n_out), (rows or n_out, n_in), "dense_kernel", group),
            VarSpec(name + "/bias", (n_out,), (n_out,), "bias", group)]


def msdn_specs():
    """MSDN variables (src/models.py:208-251) in arena order."""
    v = []
    # CoarseDense (lr 0.1, src/models.py:322-324), backward order
    # output rows stored 4070 -> 4096 (zero rows): the kernel splits into equal row slices for 2/4/8 ranks (dp.py)
    v += _dense("coarse/dense/dense_1", 4096, 4070, "CoarseDense", rows=4096)
    v += _dense("coarse/dense/dense_0", 12288, 4096, "CoarseDense")
    # CoarseConv (lr 1e-3, :319-321)
    v += _conv("coarse/conv/conv2d_4", (3, 3, 384, 256), (256, 3, 3, 384), "CoarseConv")
    v += _conv("coarse/conv/conv2d_3", (3, 3, 384, 384), (384, 3, 3, 384), "CoarseConv")
    v += _conv("coarse/conv/conv2d_2", (3, 3, 256, 384), (384, 3, 3, 256), "CoarseConv")
    # input channels stored 96 -> 128 (zero padding, as is pool0's output): 128-byte pixels, i.e. one full
    # SWIZZLE_128B row per im2col TMA pixel instead of 64-byte rows (fwd 77 -> 62 us, wgrad 125 -> 106 us)
    v += _conv("coarse/conv/conv2d_1", (5, 5, 96, 256), (256, 5, 5, 128), "CoarseConv")
    # 11x11x3 stride 4 stored as the 3x3x64 stride-1 filter over the space-to-depth(4) image (4x4 pixel blocks ->
    # 48 channels, padded to 64 = 128-byte pixels): 9 taps of 128 bytes instead of 33 of 32 bytes for TMA / UMMA
    c0 = _conv("coarse/conv/conv2d_0", (11, 11, 3, 96), (96, 3, 3, 64), "CoarseConv")
    c0[0].kind = "conv_kernel_s2d4"
    v += c0
    # FineA (lr 1e-3, :334-336): fine/third, fine/first
    v += _conv("fine/third", (5, 5, 64, 1), (1, 5, 5, 64), "FineA")
    # 9x9x3 -> 63, stride 2: the canonical (optimizer-visible) copy is the 5x5x16 -> 64 space-to-depth(2) packing;
    # the kernels read the DERIVED pool-embedded filter [256][3][3][64] (fine_first_embedded below) instead
    f1 = _conv("fine/first/conv2d", (9, 9, 3, 63), (64, 5, 5, 16), "FineA")
    f1[0].kind = "conv_kernel_s2d"
    v += f1
    # FineB (lr 0.01, :337-338)
    v += _conv("fine/second/conv2d", (5, 5, 64, 64), (64, 5, 5, 64), "FineB")
    return v


FINE_FIRST_EMBEDDED_SHAPE = (256, 3, 3, 64)


def fine_first_embedded(t: torch.Tensor) -> torch.Tensor:
    """fine/first (9x9x3 -> 63, stride 2) + max-pool 2x2 (src/models.py:241-243) as ONE stride-4 11x11 convolution
    with 4 x 64 filters, packed for the space-to-depth(4) image: [9,9,3,63] -> [256][3][3][64]."""
    return pack_s2d(embed_pool4(t, 2, 64), 4, FINE_FIRST_EMBEDDED_SHAPE)


def fine_first_index_maps(kernel_spec: "VarSpec", bias_spec: "VarSpec"):
    """(kernel map int32 [4, numel], bias map int32 [4, 64]) for a3d_gather_sum_f32 / a3d_scatter_cast_bf16."""
    def derive(ids):
        copies = []
        for g in range(4):
            big = embed_pool4(ids, 2, 64)
            sel = torch.zeros_like(big)
            sel[..., g * 64:(g + 1) * 64] = big[..., g * 64:(g + 1) * 64]
            copies.append(pack_s2d(sel, 4, FINE_FIRST_EMBEDDED_SHAPE))
        return copies
    kmap = embed_index_map(kernel_spec, derive)
    bmap = torch.full((4, bias_spec.numel), -1, dtype=torch.int32)
    n_real = bias_spec.tf_shape[0]
    for g in range(4):
        bmap[g, :n_real] = torch.arange(n_real, dtype=torch.int32) + g * 64
    return kmap, bmap


DCNF_FIRST_EMBEDDED_SHAPE = (256, 6, 2, 64)


def dcnf_first_embedded(t: torch.Tensor) -> torch.Tensor:
    """DCNF first layer (11x11x3 -> 64, stride 1) + ReLU + 2x2 max-pool (src/models.py:64-66) as ONE convolution over the
    space-to-depth(2) patch (50 x 50 cells of 16 channels, channel (2a+b)*3+c = patch pixel (2Y+a, 2X+b, c),
    a3d_extract_patches_s2d): pool-window position (dy, dx) is filter group g = 2*dy + dx, whose 11x11 filter sits at
    offset (dy, dx) inside the 12x12-pixel = 6x6-cell receptive field.  The layer reads 4 cells (64 channels, 128 bytes)
    per tap through an overlapped-pixel view, so the 6 horizontal cell taps are 2 taps of 4 cells, 4 pixels apart
    (a3d_conv_desc dil_w = 4, pix_pitch = 16): [11,11,3,64] -> [256][6][2][64], element
    [g*64 + co][tY][sv][k*16 + (2a+b)*3 + c] = w[2tY + a - dy][2(4sv + k) + b - dx][c][co]."""
    kh, kw, ci, co = t.shape
    assert (kh, kw, ci) == (11, 11, 3) and co <= 64
    big = torch.zeros(DCNF_FIRST_EMBEDDED_SHAPE, dtype=t.dtype)
    for dy in range(2):
        for dx in range(2):
            g = 2 * dy + dx
            for tY in range(6):
                for a in range(2):
                    i = 2 * tY + a - dy
                    if not 0 <= i < kh:
                        continue
                    for sv in range(2):
                        for k in range(4):
                            tX = 4 * sv + k
                            if tX >= 6:
                                continue
                            for b in range(2):
                                j = 2 * tX + b - dx
                                if not 0 <= j < kw:
                                    continue
                                ch = k * 16 + (2 * a + b) * 3
                                big[g * 64:g * 64 + co, tY, sv, ch:ch + 3] = t[i, j].t()
    return big


def dcnf_first_index_maps(kernel_spec: "VarSpec", bias_spec: "VarSpec"):
    """(kernel map int32 [4, numel], bias map int32 [4, 64]) of the embedded DCNF first layer, for a3d_gather_sum_f32 /
    a3d_scatter_cast_bf16 (same mechanism as fine_first_index_maps)."""
    def derive(ids):
        big = dcnf_first_embedded(ids)
        copies = []
        for g in range(4):
            sel = torch.zeros_like(big)
            sel[g * 64:(g + 1) * 64] = big[g * 64:(g + 1) * 64]
            copies.append(sel)
        return copies
    kmap = embed_index_map(kernel_spec, derive)
    bmap = torch.full((4, bias_spec.numel), -1, dtype=torch.int32)
    n_real = bias_spec.tf_shape[0]
    for g in range(4):
        bmap[g, :n_real] = torch.arange(n_real, dtype=torch.int32) + g * 64
    return kmap, bmap


def dcnf_specs():
    """DCNF variables (src/models.py:61-93); one SGD group (src/models.py:198)."""
    v = []
    g = "SGD"
    v += _dense("unary/unary_layers/dense_2", 16, 1, g)
    v += _dense("unary/unary_layers/dense_1", 128, 16, g)
    v += _dense("unary/unary_layers/dense", 12544, 128, g)
    v += _conv("unary/unary_layers/conv2d_4", (3, 3, 256, 256), (256, 3, 3, 256), g)
    v += _conv("unary/unary_layers/conv2d_3", (3, 3, 256, 256), (256, 3, 3, 256), g)
    v += _conv("unary/unary_layers/conv2d_2", (3, 3, 256, 256), (256, 3, 3, 256), g)
    v += _conv("unary/unary_layers/conv2d_1", (5, 5, 64, 256), (256, 5, 5, 64), g)
    v += _conv("unary/unary_layers/conv2d", (11, 11, 3, 64), (64, 11, 11, 16), g)
    v += _dense("pairwise/pairwise_layers/dense", 2, 1, "Pairwise")
    return v


class Arena:
    """Flat device buffers holding every variable: f32 master `w`, f32 grads `g`, Adam slots `m`,`v`,
    and the bf16 mirror `wb` the kernels read.  Segment starts are multiples of 64 elements: both the f32
    and the bf16 views of a segment are 16-byte aligned (TMA base alignment), and every bucket splits into
    8 rank slices that are themselves 16-byte aligned (sharded optimizer, dp.py)."""

    def __init__(self, specs, device, with_adam=True):
        off = 0
        self.specs = OrderedDict()
        self.groups = OrderedDict()
        for s in specs:
            s.offset = off
            s.size = (s.numel + 63) // 64 * 64
            off += s.size
            self.specs[s.name] = s
            lo, hi = self.groups.get(s.group, (s.offset, s.offset))
            self.groups[s.group] = (min(lo, s.offset), s.offset + s.size)
        self.total = off
        f32 = dict(dtype=torch.float32, device=device)
        self.w = torch.zeros(off, **f32)
        self.g = torch.zeros(off, **f32)
        self.m = torch.zeros(off, **f32) if with_adam else None
        self.v = torch.zeros(off, **f32) if with_adam else None
        self.wb = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.gb = None            # bf16 gradient staging for the data-parallel exchange (allocated by dp.py)
        self.masks = {}
        for s in specs:
            km = keep_mask(s)
            if km is not None:
                self.masks[s.name] = km.reshape(-1).to(device)

    def view(self, buf, name):
        s = self.specs[name]
        return buf[s.offset:s.offset + s.numel].view(s.packed_shape)

    def group_range(self, group):
        return self.groups[group]

    def num_real_params(self):
        tot = 0
        for s in self.specs.values():
            n = 1
            for d in s.tf_shape:
                n *= d
            tot += n
        return tot

    def load_tf(self, tf_params: dict, buf=None):
        """Copy TF-layout tensors (any float dtype, CPU or GPU) into the arena and refresh the mirror.
        `buf`: another arena-shaped f32 buffer (the Adam slots m / v when a checkpoint is restored) instead of w."""
        dst = self.w if buf is None else buf
        for name, s in self.specs.items():
            t = tf_params[name].detach().to("cpu", torch.float32)
            assert tuple(t.shape) == tuple(s.tf_shape), (name, t.shape, s.tf_shape)
            self.view(dst, name).copy_(pack(s, t))
        if buf is None:
            self.wb.copy_(self.w)   # load-time only; the step itself refreshes the mirror in liba3d

    def export_tf(self, buf=None):
        buf = self.w if buf is None else buf
        out = OrderedDict()
        for name, s in self.specs.items():
            out[name] = unpack(s, self.view(buf, name).detach().cpu())
        return out
