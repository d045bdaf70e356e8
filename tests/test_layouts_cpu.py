"""CPU checks of the packed layouts the CUDA path computes on (ann3depth_b200/params.py): the
space-to-depth(4) packing of coarse/conv2d_0 (src/models.py:211) and the pool-embedded form of
fine/first + max-pool (src/models.py:241-243) must be the SAME function as the TF-layout convolution the
oracle evaluates, and the index maps must fold / re-embed exactly."""
import torch
import torch.nn.functional as F

from ann3depth_b200 import params as P


def _s2d4(img):
    B, H, W, C = img.shape
    out = torch.zeros(B, H // 4, W // 4, 64, dtype=img.dtype)
    out[..., :16 * C] = img.view(B, H // 4, 4, W // 4, 4, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H // 4, W // 4, 16 * C)
    return out


def _specs():
    return {s.name: s for s in P.msdn_specs()}


def test_conv0_space_to_depth4_is_the_stride4_conv():
    g = torch.Generator().manual_seed(0)
    img = torch.rand(2, 228, 304, 3, generator=g, dtype=torch.float64)
    w = torch.rand(11, 11, 3, 96, generator=g, dtype=torch.float64) - 0.5
    ref = F.conv2d(img.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=4)
    s = _specs()["coarse/conv/conv2d_0/kernel"]
    pk = P.pack(s, w)
    assert pk.shape == (96, 3, 3, 64) and torch.equal(P.unpack(s, pk), w)
    got = F.conv2d(_s2d4(img).permute(0, 3, 1, 2), pk.permute(0, 3, 1, 2))
    assert got.shape == ref.shape == (2, 96, 55, 74)
    assert float((got - ref).abs().max()) < 1e-12


def test_fine_first_pool_embedding_is_conv_then_maxpool():
    g = torch.Generator().manual_seed(1)
    img = torch.rand(2, 228, 304, 3, generator=g, dtype=torch.float64)
    w = torch.rand(9, 9, 3, 63, generator=g, dtype=torch.float64) - 0.5
    y = F.conv2d(img.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=2)
    assert y.shape == (2, 63, 110, 148)
    pooled, arg = F.max_pool2d(y, 2, 2, return_indices=True)
    big = P.fine_first_embedded(w)
    assert big.shape == P.FINE_FIRST_EMBEDDED_SHAPE
    acc = F.conv2d(_s2d4(img).permute(0, 3, 1, 2), big.permute(0, 3, 1, 2)).view(2, 4, 64, 55, 74)
    m, gi = acc.max(1)
    assert float((m[:, :63] - pooled).abs().max()) < 1e-12
    assert float(m[:, 63].abs().max()) == 0.0                         # padding filter
    # group index g = 2a + b <-> window position (a, b) of the 2x2 pool
    a, b = gi[:, :63] // 2, gi[:, :63] % 2
    ii = torch.arange(55).view(1, 1, 55, 1)
    jj = torch.arange(74).view(1, 1, 1, 74)
    flat = (2 * ii + a) * 148 + (2 * jj + b)
    assert torch.equal(flat, arg)


def test_fine_first_index_maps_fold_and_embed():
    sp = _specs()
    ks, bs = sp["fine/first/conv2d/kernel"], sp["fine/first/conv2d/bias"]
    kmap, bmap = P.fine_first_index_maps(ks, bs)
    assert kmap.shape == (4, ks.numel) and bmap.shape == (4, 64)
    assert [int(x) for x in (kmap >= 0).sum(1)] == [9 * 9 * 3 * 63] * 4
    w = torch.rand(9, 9, 3, 63, dtype=torch.float64) - 0.5
    canon = P.pack(ks, w).reshape(-1)
    big = torch.zeros(256 * 3 * 3 * 64, dtype=torch.float64)
    for g in range(4):
        m = kmap[g] >= 0
        big[kmap[g][m].long()] = canon[m]
    assert torch.equal(big.view(P.FINE_FIRST_EMBEDDED_SHAPE), P.fine_first_embedded(w))
    # gradient fold == autograd through the embedding
    gb = torch.rand(big.numel(), dtype=torch.float64)
    fold = torch.zeros_like(canon)
    for g in range(4):
        m = kmap[g] >= 0
        fold[m] += gb[kmap[g][m].long()]
    wr = w.clone().requires_grad_(True)
    (P.fine_first_embedded(wr).reshape(-1) * gb).sum().backward()
    assert float((P.unpack(ks, fold.view(ks.packed_shape)) - wr.grad).abs().max()) < 1e-12
    assert torch.equal(bmap[:, 63], torch.full((4,), -1, dtype=torch.int32))
    assert torch.equal(bmap[2, :63], torch.arange(63, dtype=torch.int32) + 128)


def _cells(patch):
    """space-to-depth(2) of 100x100x3 patches -> flat cells [N*50*50*16] (+ zero slack), a3d_extract_patches_s2d."""
    N = patch.shape[0]
    cells = torch.zeros(N, 50, 50, 16, dtype=patch.dtype)
    for a in range(2):
        for b in range(2):
            cells[..., (2 * a + b) * 3:(2 * a + b) * 3 + 3] = patch[:, a::2, b::2, :]
    return torch.cat([cells.reshape(-1), torch.zeros(64, dtype=patch.dtype)])


def test_dcnf_first_layer_embedding_is_conv_relu_maxpool():
    """src/models.py:64-66 (11x11x3 -> 64 VALID, ReLU, 2x2 max-pool) == the 6 x 2-tap convolution (taps 4 pixels apart) of
    the embedded filter over 64-channel pixels that are 4 consecutive 16-channel cells, then the max over the 4 groups."""
    g = torch.Generator().manual_seed(0)
    w = torch.randn(11, 11, 3, 64, generator=g, dtype=torch.float64)
    bias = torch.randn(64, generator=g, dtype=torch.float64)
    patch = torch.randn(2, 100, 100, 3, generator=g, dtype=torch.float64)
    ref = F.max_pool2d(torch.relu(F.conv2d(patch.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), bias)), 2, 2)   # [2,64,45,45]
    flat = _cells(patch)
    win = flat.unfold(0, 64, 16)                      # pixel i = cells i .. i+3 (pix_pitch = 16, C = 64)
    pix = torch.arange(2 * 50 * 50).view(2, 50, 50)
    big = P.dcnf_first_embedded(w)
    assert big.shape == P.DCNF_FIRST_EMBEDDED_SHAPE
    out = torch.zeros(2, 45, 45, 256, dtype=torch.float64)
    for tY in range(6):
        for sv in range(2):
            out += win[pix[:, tY:tY + 45, 4 * sv:4 * sv + 45]] @ big[:, tY, sv, :].t()
    got = torch.relu(out.view(2, 45, 45, 4, 64).max(3).values + bias)
    assert float((got.permute(0, 3, 1, 2) - ref).abs().max()) < 1e-11


def test_dcnf_first_index_maps_fold_and_embed():
    sp = {s.name: s for s in P.dcnf_specs()}
    ks, bs = sp["unary/unary_layers/conv2d/kernel"], sp["unary/unary_layers/conv2d/bias"]
    kmap, bmap = P.dcnf_first_index_maps(ks, bs)
    assert kmap.shape == (4, ks.numel) and bmap.shape == (4, 64)
    assert [int(x) for x in (kmap >= 0).sum(1)] == [11 * 11 * 3 * 64] * 4
    w = torch.rand(11, 11, 3, 64, dtype=torch.float64) - 0.5
    canon = P.pack(ks, w).reshape(-1)
    big = torch.zeros(256 * 6 * 2 * 64, dtype=torch.float64)
    for g in range(4):
        m = kmap[g] >= 0
        big[kmap[g][m].long()] = canon[m]
    assert torch.equal(big.view(P.DCNF_FIRST_EMBEDDED_SHAPE), P.dcnf_first_embedded(w))
    gb = torch.rand(big.numel(), dtype=torch.float64)
    fold = torch.zeros_like(canon)
    for g in range(4):
        m = kmap[g] >= 0
        fold[m] += gb[kmap[g][m].long()]
    wr = w.clone().requires_grad_(True)
    (P.dcnf_first_embedded(wr).reshape(-1) * gb).sum().backward()
    assert float((P.unpack(ks, fold.view(ks.packed_shape)) - wr.grad).abs().max()) < 1e-12
    assert torch.equal(bmap[1], torch.arange(64, dtype=torch.int32) + 64)


def test_dcnf_fully_convolutional_equals_patchwise_in_float64():
    """The claim behind dcnf.py's unary="fullconv" (DESIGN.md 4.2b), checked on the CPU in float64 with the real geometry
    (240x320 image, 100x100 patches at stride 40 with a zero border of 30, src/models.py:50-83: 11x11 conv, pool, 5x5 conv,
    pool, three 3x3 convs, pool -- all VALID) and a few channels: every layer of patch (prow, pcol) is a window of the same
    layer of the zero-padded whole image, and the 7x7 input of the dense layers is the window at (5 prow, 5 pcol)."""
    from oracle import dcnf as OD
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 240, 320, 3, generator=g, dtype=torch.float64)
    chans = [3, 4, 5, 5, 5, 6]
    ks = [11, 5, 3, 3, 3]
    ws = [torch.randn(chans[i + 1], chans[i], ks[i], ks[i], generator=g, dtype=torch.float64) / ks[i] for i in range(5)]
    bs = [torch.randn(chans[i + 1], generator=g, dtype=torch.float64) * 0.1 for i in range(5)]

    def cnn(x):                                        # x [N,3,H,W] -> last pooled map
        t = F.max_pool2d(torch.relu(F.conv2d(x, ws[0], bs[0])), 2, 2)
        t = F.max_pool2d(torch.relu(F.conv2d(t, ws[1], bs[1])), 2, 2)
        t = torch.relu(F.conv2d(t, ws[2], bs[2]))
        t = torch.relu(F.conv2d(t, ws[3], bs[3]))
        return F.max_pool2d(torch.relu(F.conv2d(t, ws[4], bs[4])), 2, 2)

    patches = OD.patches(img).reshape(48, 100, 100, 3).permute(0, 3, 1, 2)            # the reference's formulation
    per_patch = cnn(patches)                                                          # [48, 6, 7, 7]
    assert per_patch.shape[-2:] == (7, 7)
    full = cnn(F.pad(img.permute(0, 3, 1, 2), (30, 30, 30, 30)))                       # once, on the padded image
    assert full.shape[-2:] == (32, 42)
    win = full.unfold(2, 7, 5).unfold(3, 7, 5)                                         # [1, C, 6, 8, 7, 7]
    assert win.shape[2:4] == (6, 8)
    gathered = win.permute(0, 2, 3, 1, 4, 5).reshape(48, chans[-1], 7, 7)
    assert float((gathered - per_patch).abs().max()) < 1e-12
