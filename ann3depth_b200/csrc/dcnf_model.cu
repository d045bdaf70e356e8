// dcnf_model.cu -- model-level entry points for DCNF (Liu et al.; `models.dcnf`, src/models.py:9-200): the whole train
// step / inference behind the C-ABI (a3d_dcnf_create / a3d_dcnf_step / a3d_dcnf_infer), the counterpart of msdn_model.cu.
// Caller-owned workspace, no allocation or host synchronisation inside a step (CUDA-graph capturable after the first,
// autotuning, call).  Mirrors ann3depth_b200/dcnf.py (unary = "fullconv") launch for launch: resize to 240x320, the unary
// CNN ONCE per zero-padded image (the reference's 100x100 patches around the 6x8 grid of 40x40 tiles are windows of that
// pass; first layer pool-fused over space-to-depth(2) cells), one 7x7x256 window per patch into the dense layers,
// pairwise colour / histogram similarities through the 2->1 dense layer, CRF negative log-likelihood with
// A = I + D - R (one CTA per graph), plain SGD (lr 0.1).  As in TF 1.3 no
// gradient reaches `pairwise_layers` (ScatterNdUpdate is not differentiable, src/models.py:138-141).
#include "common.cuh"
#include <string.h>

namespace {
constexpr int H = 240, W = 320, SP = 40, ROWS = 6, COLS = 8, NSP = ROWS * COLS;     // src/models.py:16,32-35,180-181
constexpr float SGD_LR = 0.1f, GAMMA = 1.0f;                                           // src/models.py:17,198
// fully convolutional unary network (dcnf.py, unary = "fullconv"): zero-padded image = 150 x 190 cells of 2x2 pixels; layer
// sizes per image: pooled first layer S0, conv2d_1 S1, its pool Q1, conv2d_2..4 S2..S4, last pool Q4 = 7x7 windows at stride 5
constexpr int PAD = 30, CH = (H + 2 * PAD) / 2, CW = (W + 2 * PAD) / 2;
constexpr int S0H = CH - 5, S0W = CW - 5, S1H = S0H - 4, S1W = S0W - 4, Q1H = S1H / 2, Q1W = S1W / 2;
constexpr int S2H = Q1H - 2, S2W = Q1W - 2, S3H = S2H - 2, S3W = S2W - 2, S4H = S3H - 2, S4W = S3W - 2;
constexpr int Q4H = S4H / 2, Q4W = S4W / 2;
static_assert(Q4H == (ROWS - 1) * 5 + 7 && Q4W == (COLS - 1) * 5 + 7, "7x7 windows at stride 5 tile the last pooled map");
constexpr int C0_NUMEL = 64 * 11 * 11 * 16;       // canonical first-layer filter [64][11][11][16] (3 real channels)
constexpr int EMB0_K = 256 * 6 * 2 * 64;           // its pool-embedded form [256][6][2][64]
enum { G_SGD = 0, G_PAIRWISE = 1 };

struct Seg { const char* name; int shape[4]; int ndim; int group; size_t numel, offset, size; };
// arena order = ann3depth_b200/params.py dcnf_specs()
Seg kSegs[] = {
    {"unary/unary_layers/dense_2/kernel", {1, 16, 0, 0}, 2, G_SGD},        {"unary/unary_layers/dense_2/bias", {1, 0, 0, 0}, 1, G_SGD},
    {"unary/unary_layers/dense_1/kernel", {16, 128, 0, 0}, 2, G_SGD},      {"unary/unary_layers/dense_1/bias", {16, 0, 0, 0}, 1, G_SGD},
    {"unary/unary_layers/dense/kernel", {128, 12544, 0, 0}, 2, G_SGD},     {"unary/unary_layers/dense/bias", {128, 0, 0, 0}, 1, G_SGD},
    {"unary/unary_layers/conv2d_4/kernel", {256, 3, 3, 256}, 4, G_SGD},    {"unary/unary_layers/conv2d_4/bias", {256, 0, 0, 0}, 1, G_SGD},
    {"unary/unary_layers/conv2d_3/kernel", {256, 3, 3, 256}, 4, G_SGD},    {"unary/unary_layers/conv2d_3/bias", {256, 0, 0, 0}, 1, G_SGD},
    {"unary/unary_layers/conv2d_2/kernel", {256, 3, 3, 256}, 4, G_SGD},    {"unary/unary_layers/conv2d_2/bias", {256, 0, 0, 0}, 1, G_SGD},
    {"unary/unary_layers/conv2d_1/kernel", {256, 5, 5, 64}, 4, G_SGD},     {"unary/unary_layers/conv2d_1/bias", {256, 0, 0, 0}, 1, G_SGD},
    {"unary/unary_layers/conv2d/kernel", {64, 11, 11, 16}, 4, G_SGD},      {"unary/unary_layers/conv2d/bias", {64, 0, 0, 0}, 1, G_SGD},
    {"pairwise/pairwise_layers/dense/kernel", {1, 2, 0, 0}, 2, G_PAIRWISE}, {"pairwise/pairwise_layers/dense/bias", {1, 0, 0, 0}, 1, G_PAIRWISE},
};
constexpr int NSEG = sizeof(kSegs) / sizeof(kSegs[0]);
size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

a3d_conv_desc valid_desc(int N, int Hh, int Ww, int C, int K, int R) {
  a3d_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.N = N; d.H = Hh; d.W = Ww; d.C = C; d.K = K; d.R = d.S = R; d.stride_h = d.stride_w = 1;
  d.P = Hh - R + 1; d.Q = Ww - R + 1; d.ldy = K; d.impl = A3D_IMPL_AUTO;
  return d;
}
}  // namespace

struct a3d_dcnf {
  a3d_ctx* ctx;
  int B, NP, inH, inW, dH, dW, train, naive;
  long long global_step;
  Seg seg[NSEG];
  size_t total, sgd_lo, sgd_hi;
  float *w, *g; uint16_t* wb;
  float *im, *dp, *c1, *c4, *z, *sims, *r, *y, *ystar, *nll, *logdet, *loss, *output, *dz, *dense_acc, *pair_ws;
  uint16_t *cells, *wbig0, *p0, *p1, *c2, *c3, *p4, *xd, *h0, *h1, *g_big0, *g_xd;
  float* g_wbig0;
  int32_t *emb_k, *emb_b;
  uint8_t *i0, *i1, *i4;
  int32_t *pl, *pr, *status;
  uint16_t *g_z, *g_h1a, *g_h1, *g_h0a, *g_h0, *g_p4, *g_c4, *g_c3, *g_c2, *g_p1, *g_c1, *g_p0;
  void* scratch; size_t scratch_bytes;
  a3d_conv_desc d0, d0w, d1, d2, d3, d4;
};

namespace {
size_t carve(a3d_dcnf* n, uint8_t* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> void* {
    void* p = base ? base + off : nullptr;
    off += al256(bytes);
    return p;
  };
  const size_t B = n->B, NP = n->NP, T = n->total;
  n->w = (float*)take(T * 4); n->g = (float*)take(T * 4); n->wb = (uint16_t*)take(T * 2);
  n->im = (float*)take(B * H * W * 3 * 4);
  n->dp = (float*)take(B * H * W * 4);
  n->cells = (uint16_t*)take(B * CH * CW * 16 * 2 + 256);          // + slack: the overlapped view reads past the last row
  n->wbig0 = (uint16_t*)take((size_t)EMB0_K * 2);
  n->emb_k = (int32_t*)take((size_t)4 * C0_NUMEL * 4);
  n->emb_b = (int32_t*)take(4 * 64 * 4);
  n->p0 = (uint16_t*)take(B * S0H * S0W * 64 * 2);
  n->i0 = (uint8_t*)take(B * S0H * S0W * 64);
  n->c1 = (float*)take(B * S1H * S1W * 256 * 4);
  n->p1 = (uint16_t*)take(B * Q1H * Q1W * 256 * 2);
  n->i1 = (uint8_t*)take(B * Q1H * Q1W * 256);
  n->c2 = (uint16_t*)take(B * S2H * S2W * 256 * 2);
  n->c3 = (uint16_t*)take(B * S3H * S3W * 256 * 2);
  n->c4 = (float*)take(B * S4H * S4W * 256 * 4);
  n->p4 = (uint16_t*)take(B * Q4H * Q4W * 256 * 2);
  n->i4 = (uint8_t*)take(B * Q4H * Q4W * 256);
  n->xd = (uint16_t*)take(NP * 12544 * 2);
  n->h0 = (uint16_t*)take(NP * 128 * 2);
  n->h1 = (uint16_t*)take(NP * 16 * 2);
  n->z = (float*)take(NP * 4);
  n->sims = (float*)take(B * NSP * 2 * 4);
  n->r = (float*)take(B * NSP * 4);
  n->y = (float*)take(B * NSP * 4);
  n->ystar = (float*)take(B * NSP * 4);
  n->nll = (float*)take(B * 4);
  n->logdet = (float*)take(B * 4);
  n->status = (int32_t*)take(B * 4);
  n->loss = (float*)take(4);
  n->output = (float*)take(B * H * W * 4);
  n->pl = (int32_t*)take(NSP * 4);
  n->pr = (int32_t*)take(NSP * 4);
  n->dense_acc = (float*)take(NP * 12544 * 4);
  n->pair_ws = (float*)take(a3d_pairwise_ws_bytes((int)B, H, W));
  if (n->train) {
    n->dz = (float*)take(B * NSP * 4);
    n->g_z = (uint16_t*)take(NP * 2);
    n->g_h1a = (uint16_t*)take(NP * 16 * 2); n->g_h1 = (uint16_t*)take(NP * 16 * 2);
    n->g_h0a = (uint16_t*)take(NP * 128 * 2); n->g_h0 = (uint16_t*)take(NP * 128 * 2);
    n->g_xd = (uint16_t*)take(NP * 12544 * 2);
    n->g_p4 = (uint16_t*)take(B * Q4H * Q4W * 256 * 2);
    n->g_c4 = (uint16_t*)take(B * S4H * S4W * 256 * 2);
    n->g_c3 = (uint16_t*)take(B * S3H * S3W * 256 * 2);
    n->g_c2 = (uint16_t*)take(B * S2H * S2W * 256 * 2);
    n->g_p1 = (uint16_t*)take(B * Q1H * Q1W * 256 * 2);
    n->g_c1 = (uint16_t*)take(B * S1H * S1W * 256 * 2);
    n->g_p0 = (uint16_t*)take(B * S0H * S0W * 64 * 2);
    n->g_big0 = (uint16_t*)take(B * S0H * S0W * 256 * 2);
    n->g_wbig0 = (float*)take(((size_t)EMB0_K + 256) * 4);
  }
  size_t sc = 256;
  const a3d_conv_desc* ds[] = {&n->d0w, &n->d1, &n->d2, &n->d3, &n->d4};
  for (const a3d_conv_desc* d : ds)
    for (int op = A3D_OP_FWD; op <= A3D_OP_WGRAD; ++op) {
      size_t b = a3d_conv2d_ws_bytes(n->ctx, d, op);
      if (b > sc) sc = b;
    }
  n->scratch_bytes = al256(sc);
  n->scratch = take(n->scratch_bytes);
  return off;
}

void init_layout(a3d_dcnf* n, a3d_ctx* ctx, int batch, int in_h, int in_w, int dh, int dw, int train) {
  memset(n, 0, sizeof(*n));
  n->ctx = ctx; n->B = batch; n->NP = batch * NSP; n->inH = in_h; n->inW = in_w; n->dH = dh; n->dW = dw; n->train = train;
  n->naive = 1;                                           // the reference's loss form (src/models.py:166-171)
  size_t off = 0;
  n->sgd_lo = (size_t)-1; n->sgd_hi = 0;
  for (int i = 0; i < NSEG; ++i) {
    Seg s = kSegs[i];
    s.numel = 1;
    for (int k = 0; k < s.ndim; ++k) s.numel *= (size_t)s.shape[k];
    s.offset = off;
    s.size = (s.numel + 63) / 64 * 64;
    off += s.size;
    n->seg[i] = s;
    if (s.group == G_SGD) {
      if (s.offset < n->sgd_lo) n->sgd_lo = s.offset;
      if (s.offset + s.size > n->sgd_hi) n->sgd_hi = s.offset + s.size;
    }
  }
  n->total = off;
  // src/models.py:64-66 (11x11x3 -> 64, ReLU, 2x2 max-pool) as one pool-fused 6x6-cell convolution over the
  // space-to-depth(2) image, read through the overlapped-pixel view (a3d_image_cells_s2d, dcnf.py)
  n->d0 = valid_desc(batch, CH, CW, 64, 256, 6);
  n->d0.S = 2; n->d0.P = S0H; n->d0.Q = S0W; n->d0.ldy = 64; n->d0.dil_w = 4; n->d0.pix_pitch = 16;
  n->d0w = n->d0;
  n->d0w.ldy = 256;
  n->d1 = valid_desc(batch, S0H, S0W, 64, 256, 5);        // :67
  n->d2 = valid_desc(batch, Q1H, Q1W, 256, 256, 3);       // :69
  n->d3 = valid_desc(batch, S2H, S2W, 256, 256, 3);       // :71
  n->d4 = valid_desc(batch, S3H, S3W, 256, 256, 3);       // :72
}

size_t seg_off(const a3d_dcnf* n, const char* name) {
  for (int i = 0; i < NSEG; ++i)
    if (!strcmp(n->seg[i].name, name)) return n->seg[i].offset;
  return 0;
}
#define CK(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)
#define U "unary/unary_layers/"

int forward(a3d_dcnf* n, const float* images, const float* depths, void* st) {
  a3d_ctx* c = n->ctx;
  const int B = n->B, NP = n->NP;
  auto Wb = [&](const char* nm) { return n->wb + seg_off(n, nm); };
  auto Wf = [&](const char* nm) { return n->w + seg_off(n, nm); };
  CK(a3d_resize_bilinear_tf1(c, images, B, n->inH, n->inW, 3, n->im, H, W, 3, A3D_F32, st));
  if (depths) CK(a3d_resize_bilinear_tf1(c, depths, B, n->dH, n->dW, 1, n->dp, H, W, 1, A3D_F32, st));
  // unary part (src/models.py:61-89) on all B*48 patches at once
  CK(a3d_image_cells_s2d(c, n->im, B, H, W, PAD, n->cells, st));
  CK(a3d_conv2d_pool4_fwd(c, &n->d0, n->cells, n->wbig0, Wf(U "conv2d/bias"), n->p0, n->i0, A3D_EPI_RELU, nullptr, 0, st));
  CK(a3d_conv2d_fwd(c, &n->d1, n->p0, Wb(U "conv2d_1/kernel"), Wf(U "conv2d_1/bias"), n->c1, A3D_F32, A3D_EPI_RELU, n->scratch,
                    n->scratch_bytes, st));
  CK(a3d_maxpool2x2_fwd_f32(c, n->c1, B, S1H, S1W, 256, n->p1, 256, n->i1, st));
  CK(a3d_conv2d_fwd(c, &n->d2, n->p1, Wb(U "conv2d_2/kernel"), Wf(U "conv2d_2/bias"), n->c2, A3D_BF16, A3D_EPI_RELU, n->scratch,
                    n->scratch_bytes, st));
  CK(a3d_conv2d_fwd(c, &n->d3, n->c2, Wb(U "conv2d_3/kernel"), Wf(U "conv2d_3/bias"), n->c3, A3D_BF16, A3D_EPI_RELU, n->scratch,
                    n->scratch_bytes, st));
  CK(a3d_conv2d_fwd(c, &n->d4, n->c3, Wb(U "conv2d_4/kernel"), Wf(U "conv2d_4/bias"), n->c4, A3D_F32, A3D_EPI_RELU, n->scratch,
                    n->scratch_bytes, st));
  CK(a3d_maxpool2x2_fwd_f32(c, n->c4, B, S4H, S4W, 256, n->p4, 256, n->i4, st));
  // patch (prow, pcol) reads the 7x7x256 window at (5 prow, 5 pcol) of the last pooled map
  CK(a3d_window_gather(c, n->p4, B, Q4H, Q4W, 256, ROWS, COLS, 7, 5, n->xd, st));
  // the 768-row dense layer goes in chunks of 256 rows inside a3d_dense_fwd; the two tiny ones run on the CUDA cores
  CK(a3d_dense_fwd(c, n->xd, 12544, Wb(U "dense/kernel"), Wf(U "dense/bias"), nullptr, 0.f, n->h0, A3D_BF16, n->dense_acc, NP, 128,
                   12544, A3D_EPI_RELU, A3D_IMPL_AUTO, st));
  CK(a3d_dense_fwd(c, n->h0, 128, Wb(U "dense_1/kernel"), Wf(U "dense_1/bias"), nullptr, 0.f, n->h1, A3D_BF16, n->dense_acc, NP, 16,
                   128, A3D_EPI_SIGMOID, A3D_IMPL_SIMT, st));
  CK(a3d_dense_fwd(c, n->h1, 16, Wb(U "dense_2/kernel"), Wf(U "dense_2/bias"), nullptr, 0.f, n->z, A3D_F32, n->dense_acc, NP, 1, 16, 0,
                   A3D_IMPL_SIMT, st));
  // pairwise part (src/models.py:108-127)
  CK(a3d_pairwise_features(c, n->im, B, H, W, n->pl, n->pr, NSP, GAMMA, n->pair_ws, n->sims, st));
  CK(a3d_pairwise_dense(c, n->sims, Wf("pairwise/pairwise_layers/dense/kernel"), Wf("pairwise/pairwise_layers/dense/bias"), n->r,
                        (size_t)B * NSP, st));
  if (depths) {
    // loss part (src/models.py:129-177)
    CK(a3d_tile_means(c, n->dp, B, H, W, n->y, st));
    CK(a3d_crf_fwd_bwd(c, n->z, n->y, n->r, n->pl, n->pr, B, NSP, NSP, 1.0f / B, n->naive, n->ystar, n->nll, n->logdet,
                       n->train ? n->dz : nullptr, nullptr, n->status, st));
    CK(a3d_mean_f32(c, n->nll, B, n->loss, st));
  }
  // output (src/models.py:187-191): the unary prediction, upsampled
  return a3d_resize_bilinear_tf1(c, n->z, B, ROWS, COLS, 1, n->output, H, W, 1, A3D_F32, st);
}

int backward(a3d_dcnf* n, void* st) {
  a3d_ctx* c = n->ctx;
  const int NP = n->NP;
  auto Wb = [&](const char* nm) { return n->wb + seg_off(n, nm); };
  auto G = [&](const char* nm) { return n->g + seg_off(n, nm); };
  const int S = A3D_IMPL_SIMT;
  CK(a3d_scale_cast_bf16(c, n->dz, n->g_z, (size_t)NP, 1.0f, st));
  CK(a3d_dense_wgrad(c, n->h1, 16, n->g_z, 1, G(U "dense_2/kernel"), G(U "dense_2/bias"), NP, 1, 16, S, st));
  CK(a3d_dense_dgrad(c, n->g_z, 1, Wb(U "dense_2/kernel"), n->g_h1a, n->dense_acc, NP, 1, 16, S, st));
  CK(a3d_dense_epilogue_bwd(c, n->g_h1a, n->h1, nullptr, 0.f, n->g_h1, (size_t)NP * 16, A3D_EPI_SIGMOID, st));
  CK(a3d_dense_wgrad(c, n->h0, 128, n->g_h1, 16, G(U "dense_1/kernel"), G(U "dense_1/bias"), NP, 16, 128, S, st));
  CK(a3d_dense_dgrad(c, n->g_h1, 16, Wb(U "dense_1/kernel"), n->g_h0a, n->dense_acc, NP, 16, 128, S, st));
  CK(a3d_dense_epilogue_bwd(c, n->g_h0a, n->h0, nullptr, 0.f, n->g_h0, (size_t)NP * 128, A3D_EPI_RELU, st));
  CK(a3d_dense_wgrad(c, n->xd, 12544, n->g_h0, 128, G(U "dense/kernel"), G(U "dense/bias"), NP, 128, 12544, A3D_IMPL_AUTO, st));
  CK(a3d_dense_dgrad(c, n->g_h0, 128, Wb(U "dense/kernel"), n->g_xd, n->dense_acc, NP, 128, 12544, A3D_IMPL_AUTO, st));
  CK(a3d_window_scatter_sum(c, n->g_xd, n->B, Q4H, Q4W, 256, ROWS, COLS, 7, 5, n->g_p4, st));
  CK(a3d_maxpool2x2_idx_bwd(c, n->i4, n->g_p4, 256, n->B, S4H, S4W, 256, n->g_c4, st));
  CK(a3d_conv2d_wgrad(c, &n->d4, n->c3, n->g_c4, G(U "conv2d_4/kernel"), G(U "conv2d_4/bias"), n->scratch, n->scratch_bytes, st));
  CK(a3d_conv2d_dgrad(c, &n->d4, n->g_c4, Wb(U "conv2d_4/kernel"), n->g_c3, n->c3, n->scratch, n->scratch_bytes, st));
  CK(a3d_conv2d_wgrad(c, &n->d3, n->c2, n->g_c3, G(U "conv2d_3/kernel"), G(U "conv2d_3/bias"), n->scratch, n->scratch_bytes, st));
  CK(a3d_conv2d_dgrad(c, &n->d3, n->g_c3, Wb(U "conv2d_3/kernel"), n->g_c2, n->c2, n->scratch, n->scratch_bytes, st));
  CK(a3d_conv2d_wgrad(c, &n->d2, n->p1, n->g_c2, G(U "conv2d_2/kernel"), G(U "conv2d_2/bias"), n->scratch, n->scratch_bytes, st));
  CK(a3d_conv2d_dgrad(c, &n->d2, n->g_c2, Wb(U "conv2d_2/kernel"), n->g_p1, nullptr, n->scratch, n->scratch_bytes, st));
  CK(a3d_maxpool2x2_idx_bwd(c, n->i1, n->g_p1, 256, n->B, S1H, S1W, 256, n->g_c1, st));
  CK(a3d_conv2d_wgrad(c, &n->d1, n->p0, n->g_c1, G(U "conv2d_1/kernel"), G(U "conv2d_1/bias"), n->scratch, n->scratch_bytes, st));
  CK(a3d_conv2d_dgrad(c, &n->d1, n->g_c1, Wb(U "conv2d_1/kernel"), n->g_p0, nullptr, n->scratch, n->scratch_bytes, st));
  // first layer: MaxPoolGrad + ReluGrad onto the 4 x 64 GEMM columns, weight gradient of the embedded filter, then its
  // four copies (and the four bias groups) folded into the canonical variable (the 13 padding channels receive 0)
  CK(a3d_pool4_bwd(c, n->g_p0, 64, n->p0, 64, n->i0, n->g_big0, (size_t)n->B * S0H * S0W, st));
  CK(a3d_conv2d_wgrad(c, &n->d0w, n->cells, n->g_big0, n->g_wbig0, n->g_wbig0 + EMB0_K, n->scratch, n->scratch_bytes, st));
  CK(a3d_gather_sum_f32(c, n->g_wbig0, n->emb_k, 4, C0_NUMEL, G(U "conv2d/kernel"), st));
  return a3d_gather_sum_f32(c, n->g_wbig0 + EMB0_K, n->emb_b, 4, 64, G(U "conv2d/bias"), st);
}
}  // namespace

extern "C" size_t a3d_dcnf_workspace_bytes(a3d_ctx* ctx, int batch, int in_h, int in_w, int depth_h, int depth_w, int train) {
  if (!ctx || batch <= 0) return 0;
  a3d_dcnf n;
  init_layout(&n, ctx, batch, in_h, in_w, depth_h, depth_w, train);
  return carve(&n, nullptr) + 256;
}

extern "C" int a3d_dcnf_create(a3d_ctx* ctx, int batch, int in_h, int in_w, int depth_h, int depth_w, int train, void* workspace,
                               size_t workspace_bytes, void* stream, a3d_dcnf** out) {
  A3D_REQUIRE(ctx && workspace && out && batch > 0 && in_h > 0 && in_w > 0 && depth_h > 0 && depth_w > 0, "dcnf_create: bad argument");
  a3d_dcnf* n = new a3d_dcnf();
  init_layout(n, ctx, batch, in_h, in_w, depth_h, depth_w, train);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const size_t need = carve(n, base);
  if (base + need > reinterpret_cast<uint8_t*>(workspace) + workspace_bytes) {
    delete n;
    a3d_set_error("dcnf_create: workspace of %zu bytes is too small (a3d_dcnf_workspace_bytes)", workspace_bytes);
    return A3D_EINVAL;
  }
  cudaStream_t st = as_stream(stream);
  A3D_CHECK_CUDA(cudaMemsetAsync(base, 0, need, st));
  // the static pair graph (src/models.py:20-30): interior checkerboard centres x 4 neighbours = 48 directed pairs
  int32_t pl[NSP], pr[NSP];
  int np = 0;
  for (int row = 1; row < ROWS - 1; ++row)
    for (int col = 2 - (row & 1); col < COLS - 1; col += 2) {
      const int pixel = row * COLS + col;
      const int add[4] = {-COLS, COLS, -1, 1};
      for (int k = 0; k < 4; ++k) { pl[np] = pixel; pr[np] = pixel + add[k]; ++np; }
    }
  if (np != NSP) { delete n; a3d_set_error("dcnf_create: pair graph has %d pairs, expected %d", np, NSP); return A3D_EINVAL; }
  // index maps canonical [64][11][11][16] -> the four embedded copies inside [256][6][2][64] (params.dcnf_first_embedded)
  int32_t* mk = new int32_t[4 * C0_NUMEL + 4 * 64];
  for (int i = 0; i < 4 * C0_NUMEL; ++i) mk[i] = -1;
  for (int g = 0; g < 4; ++g) {
    const int dy = g >> 1, dx = g & 1;
    for (int co = 0; co < 64; ++co)
      for (int tY = 0; tY < 6; ++tY)
        for (int a = 0; a < 2; ++a) {
          const int i = 2 * tY + a - dy;
          if (i < 0 || i >= 11) continue;
          for (int tX = 0; tX < 6; ++tX)
            for (int b = 0; b < 2; ++b) {
              const int j = 2 * tX + b - dx;
              if (j < 0 || j >= 11) continue;
              for (int ch = 0; ch < 3; ++ch) {
                const int canon = ((co * 11 + i) * 11 + j) * 16 + ch;
                const int derived = (((g * 64 + co) * 6 + tY) * 2 + tX / 4) * 64 + (tX % 4) * 16 + (2 * a + b) * 3 + ch;
                mk[g * C0_NUMEL + canon] = derived;
              }
            }
        }
    for (int co = 0; co < 64; ++co) mk[4 * C0_NUMEL + g * 64 + co] = g * 64 + co;
  }
  cudaError_t e1 = cudaMemcpyAsync(n->pl, pl, sizeof(pl), cudaMemcpyHostToDevice, st);
  cudaError_t e2 = cudaMemcpyAsync(n->pr, pr, sizeof(pr), cudaMemcpyHostToDevice, st);
  cudaError_t e3 = cudaMemcpyAsync(n->emb_k, mk, (size_t)4 * C0_NUMEL * 4, cudaMemcpyHostToDevice, st);
  if (e3 == cudaSuccess) e3 = cudaMemcpyAsync(n->emb_b, mk + 4 * C0_NUMEL, 4 * 64 * 4, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  delete[] mk;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { delete n; a3d_set_error("dcnf_create: upload failed"); return A3D_EINVAL; }
  *out = n;
  return 0;
}

extern "C" int a3d_dcnf_destroy(a3d_dcnf* n) { delete n; return 0; }

// naive_loss = 1: the reference's exp(-E)/Z form (saturates at 16.118); 0: the numerically stable closed form
extern "C" int a3d_dcnf_configure(a3d_dcnf* n, int naive_loss) {
  A3D_REQUIRE(n, "dcnf_configure: null net");
  n->naive = naive_loss ? 1 : 0;
  return 0;
}

extern "C" int a3d_dcnf_segment(const a3d_dcnf* n, int index, const char** name, size_t* offset, size_t* numel, int shape[4]) {
  if (!n) return 0;
  if (index < 0 || index >= NSEG) return NSEG;
  const Seg& s = n->seg[index];
  if (name) *name = s.name;
  if (offset) *offset = s.offset;
  if (numel) *numel = s.numel;
  if (shape) for (int k = 0; k < 4; ++k) shape[k] = k < s.ndim ? s.shape[k] : 0;
  return NSEG;
}

extern "C" int a3d_dcnf_arena(a3d_dcnf* n, float** w, float** g, uint16_t** w_bf16, size_t* total) {
  A3D_REQUIRE(n, "dcnf_arena: null net");
  if (w) *w = n->w;
  if (g) *g = n->g;
  if (w_bf16) *w_bf16 = n->wb;
  if (total) *total = n->total;
  return 0;
}

extern "C" int a3d_dcnf_sync_weights(a3d_dcnf* n, void* stream) {
  A3D_REQUIRE(n, "dcnf_sync_weights: null net");
  CK(a3d_cast_f32_bf16(n->ctx, n->w, n->wb, n->total, stream));
  return a3d_scatter_cast_bf16(n->ctx, n->w + seg_off(n, U "conv2d/kernel"), n->emb_k, 4, C0_NUMEL, n->wbig0, stream);
}

extern "C" long long a3d_dcnf_global_step(const a3d_dcnf* n) { return n ? n->global_step : -1; }

// One `session.run(model_op)` of models.dcnf: forward, CRF loss, backward through the unary CNN, SGD (src/models.py:198-200:
// minimize(loss, global_step)).  images f32 [B,in_h,in_w,3], depths f32 [B,depth_h,depth_w,1] (device).  loss (nullable,
// device, 1 float) receives the mean negative log-likelihood.  Kernel launches only: capturable.
extern "C" int a3d_dcnf_step(a3d_dcnf* n, const float* images, const float* depths, float* loss, void* stream) {
  A3D_REQUIRE(n && n->train && images && depths, "dcnf_step: needs a net created with train = 1, images and depths");
  CK(forward(n, images, depths, stream));
  CK(backward(n, stream));
  CK(a3d_sgd(n->ctx, n->w + n->sgd_lo, n->g + n->sgd_lo, n->wb + n->sgd_lo, n->sgd_hi - n->sgd_lo, SGD_LR, 1.0f, stream));
  CK(a3d_scatter_cast_bf16(n->ctx, n->w + seg_off(n, U "conv2d/kernel"), n->emb_k, 4, C0_NUMEL, n->wbig0, stream));
  n->global_step += 1;
  if (loss) A3D_CHECK_CUDA(cudaMemcpyAsync(loss, n->loss, sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
  return 0;
}

// Inference: the unary prediction z [B,48] and its bilinear upsampling output [B,240,320] (src/models.py:187-191), plus
// (optional) the pairwise potentials r [B,48].  All device pointers, nullable.
extern "C" int a3d_dcnf_infer(a3d_dcnf* n, const float* images, float* output, float* z, float* r, void* stream) {
  A3D_REQUIRE(n && images, "dcnf_infer: bad argument");
  CK(forward(n, images, nullptr, stream));
  cudaStream_t st = as_stream(stream);
  if (output) A3D_CHECK_CUDA(cudaMemcpyAsync(output, n->output, (size_t)n->B * H * W * 4, cudaMemcpyDeviceToDevice, st));
  if (z) A3D_CHECK_CUDA(cudaMemcpyAsync(z, n->z, (size_t)n->NP * 4, cudaMemcpyDeviceToDevice, st));
  if (r) A3D_CHECK_CUDA(cudaMemcpyAsync(r, n->r, (size_t)n->B * NSP * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// device pointers into the net's state for hosts that want more than the loss: CRF MAP estimate y* = A^-1 z [B,48],
// per-graph status (0 = SPD), the loss scalar
extern "C" int a3d_dcnf_state(a3d_dcnf* n, const float** ystar, const int32_t** status, const float** loss) {
  A3D_REQUIRE(n, "dcnf_state: null net");
  if (ystar) *ystar = n->ystar;
  if (status) *status = n->status;
  if (loss) *loss = n->loss;
  return 0;
}
