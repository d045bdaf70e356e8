// f32_ops.cu -- memory-bound kernels of the TF32 precision mode (models.msdn(..., dtype="tf32")): every activation and
// activation gradient is stored in float32 (the reference's own storage type, src/models.py:211-251), the contractions
// run on tcgen05.mma kind::tf32 (tc_gemm.cu).  Same operations as their bf16-storage counterparts in elementwise.cu /
// conv.cu; plain float4-vectorised grid-stride kernels (the mode exists for precision, its step time is reported beside
// the bf16 mode's, not tuned to it).
#include "common.cuh"

namespace {
inline int grid_f32(a3d_ctx* ctx, size_t work, int block = 256) {
  size_t b = (work + block - 1) / block;
  const size_t cap = (size_t)ctx->sm_count * 16;
  return (int)(b < 1 ? 1 : b > cap ? cap : b);
}
}  // namespace

// ---- resize (TF1 legacy bilinear) + space-to-depth(s), float32 output [B, OH/s, OW/s, dstC >= s*s*C] --------------------
__global__ void resize_s2d_f32_kernel(const float* __restrict__ src, int B, int H, int W, int C, float* __restrict__ dst,
                                      int OHs, int OWs, int s, int dstC, float sy, float sx) {
  const size_t total = (size_t)B * OHs * OWs * dstC;
  const int real = s * s * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % dstC);
    size_t t = i / dstC;
    const int X = (int)(t % OWs);
    t /= OWs;
    const int Y = (int)(t % OHs);
    const int b = (int)(t / OHs);
    float r = 0.f;
    if (k < real) {
      const int pix = k / C, c = k - pix * C;
      const int dy = pix / s, dx = pix - dy * s;
      const int oy = Y * s + dy, ox = X * s + dx;
      const float fy = oy * sy, fx = ox * sx;
      const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
      const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
      const float ly = fy - y0, lx = fx - x0;
      const float* r0 = src + ((size_t)b * H + y0) * W * C;
      const float* r1 = src + ((size_t)b * H + y1) * W * C;
      const float tl = __ldg(r0 + (size_t)x0 * C + c), tr = __ldg(r0 + (size_t)x1 * C + c);
      const float bl = __ldg(r1 + (size_t)x0 * C + c), br = __ldg(r1 + (size_t)x1 * C + c);
      const float top = tl + (tr - tl) * lx;
      const float bot = bl + (br - bl) * lx;
      r = top + (bot - top) * ly;
    }
    dst[i] = r;
  }
}

extern "C" int a3d_resize_bilinear_tf1_s2d_f32(a3d_ctx* ctx, const float* src, int B, int H, int W, int C, float* dst,
                                               int OH, int OW, int s, int dstC, void* stream) {
  A3D_REQUIRE(ctx && src && dst && s > 0 && OH % s == 0 && OW % s == 0 && dstC >= s * s * C, "resize_s2d_f32: bad argument");
  const size_t total = (size_t)B * (OH / s) * (OW / s) * dstC;
  // float32 scale as in TF's CalculateResizeScale (and the bf16-output kernel)
  const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
  resize_s2d_f32_kernel<<<grid_f32(ctx, total), 256, 0, as_stream(stream)>>>(src, B, H, W, C, dst, OH / s, OW / s, s, dstC,
                                                                             sy, sx);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- 2x2/2 max-pool, float32 in and out, with the 1-byte routing record (first arg-max, TF MaxPoolGrad + ReluGrad) ----
__global__ void maxpool2x2_f32_kernel(const float* __restrict__ x, int N, int H, int W, int C, float* __restrict__ y, int ldy,
                                      uint8_t* __restrict__ idx) {
  const int OH = H / 2, OW = W / 2;
  const size_t total = (size_t)N * OH * OW * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n, oh, ow, c;
    split_nhwc(i, OH, OW, C, n, oh, ow, c);
    const float* p = x + (((size_t)n * H + 2 * oh) * W + 2 * ow) * C + c;
    float m = __ldg(p);
    int g = 0;
    const float a1 = __ldg(p + C), a2 = __ldg(p + (size_t)W * C), a3 = __ldg(p + (size_t)W * C + C);
    if (a1 > m) { m = a1; g = 1; }
    if (a2 > m) { m = a2; g = 2; }
    if (a3 > m) { m = a3; g = 3; }
    y[(((size_t)n * OH + oh) * OW + ow) * ldy + c] = m;
    // routing record: the pooled tensor is a ReLU output everywhere in these models, and ReluGrad is folded into the
    // record (4 = "no route": the window's maximum is not positive), exactly as a3d_maxpool2x2_fwd_f32 does
    if (idx) idx[i] = (uint8_t)(m > 0.f ? g : 4);
  }
}

extern "C" int a3d_maxpool2x2_f32(a3d_ctx* ctx, const float* x, int N, int H, int W, int C, float* y, int ldy, uint8_t* idx,
                                  void* stream) {
  A3D_REQUIRE(ctx && x && y && ldy >= C, "maxpool2x2_f32: bad argument");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * C;
  maxpool2x2_f32_kernel<<<grid_f32(ctx, total), 256, 0, as_stream(stream)>>>(x, N, H, W, C, y, ldy, idx);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// dx[n,h,w,c] = dy[n,h/2,w/2,c] where (h%2)*2 + w%2 is the recorded arg-max, 0 elsewhere (and in a dropped odd edge)
__global__ void maxpool2x2_idx_bwd_f32_kernel(const uint8_t* __restrict__ idx, const float* __restrict__ dy, int lddy, int N,
                                              int H, int W, int C, float* __restrict__ dx) {
  const int OH = H / 2, OW = W / 2;
  const size_t total = (size_t)N * H * W * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n, h, w, c;
    split_nhwc(i, H, W, C, n, h, w, c);
    const int oh = h >> 1, ow = w >> 1;
    float v = 0.f;
    if (oh < OH && ow < OW) {
      const size_t o = ((size_t)n * OH + oh) * OW + ow;
      if (idx[o * C + c] == (uint8_t)((h & 1) * 2 + (w & 1))) v = __ldg(dy + o * lddy + c);
    }
    dx[i] = v;
  }
}

extern "C" int a3d_maxpool2x2_idx_bwd_f32(a3d_ctx* ctx, const uint8_t* idx, const float* dy, int lddy, int N, int H, int W,
                                          int C, float* dx, void* stream) {
  A3D_REQUIRE(ctx && idx && dy && dx && lddy >= C, "maxpool2x2_idx_bwd_f32: bad argument");
  maxpool2x2_idx_bwd_f32_kernel<<<grid_f32(ctx, (size_t)N * H * W * C), 256, 0, as_stream(stream)>>>(idx, dy, lddy, N, H, W,
                                                                                                    C, dx);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- ReluGrad / DropoutGrad / SigmoidGrad on float32 (dx may alias dy) ------------------------------------------------
__global__ void act_bwd_f32_kernel(const float* __restrict__ g_post, int ldg, const float* __restrict__ y,
                                   const uint8_t* __restrict__ mask, float scale, float* __restrict__ g_pre, size_t rows,
                                   int C, unsigned flags) {
  const size_t total = rows * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / C;
    const int c = (int)(i - r * C);
    float g = g_post[r * ldg + c];
    if (mask) g = mask[i] ? g * scale : 0.f;
    const float yv = y[i];
    if (flags & A3D_EPI_RELU) g = yv > 0.f ? g : 0.f;
    if (flags & A3D_EPI_SIGMOID) g = g * yv * (1.f - yv);
    g_pre[i] = g;
  }
}

extern "C" int a3d_act_bwd_f32(a3d_ctx* ctx, const float* g_post, int ldg, const float* y, const uint8_t* keep_mask,
                               float drop_rate, float* g_pre, size_t rows, int C, unsigned flags, void* stream) {
  A3D_REQUIRE(ctx && g_post && y && g_pre && ldg >= C, "act_bwd_f32: bad argument");
  act_bwd_f32_kernel<<<grid_f32(ctx, rows * C), 256, 0, as_stream(stream)>>>(g_post, ldg, y, keep_mask,
                                                                             1.f / (1.f - drop_rate), g_pre, rows, C, flags);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- dst[row*ld + ch] = src[row] (the coarse map as channel 63 of the fine stack's concat, src/models.py:246) ----------
__global__ void scatter_channel_f32_kernel(const float* __restrict__ s, float* __restrict__ d, size_t rows, int ld, int ch) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x)
    d[i * ld + ch] = s[i];
}
extern "C" int a3d_scatter_channel_f32(a3d_ctx* ctx, const float* src, float* dst, size_t rows, int ld, int ch, void* stream) {
  A3D_REQUIRE(ctx && src && dst && ch < ld, "scatter_channel_f32: bad argument");
  scatter_channel_f32_kernel<<<grid_f32(ctx, rows), 256, 0, as_stream(stream)>>>(src, dst, rows, ld, ch);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- pool over the four 64-column groups of the pool-embedded fine/first GEMM (see a3d_conv2d_pool4_fwd) ---------------
__global__ void pool4_reduce_f32_kernel(const float* __restrict__ acc, const float* __restrict__ bias, float* __restrict__ y,
                                        int ldy, uint8_t* __restrict__ idx, size_t rows, unsigned flags) {
  const size_t total = rows * 64;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i >> 6;
    const int c = (int)(i & 63);
    const float* a = acc + row * 256 + c;
    float m = a[0];
    int g = 0;
    if (a[64] > m) { m = a[64]; g = 1; }
    if (a[128] > m) { m = a[128]; g = 2; }
    if (a[192] > m) { m = a[192]; g = 3; }
    if (bias) m += bias[c];
    if (flags & A3D_EPI_RELU) m = fmaxf(m, 0.f);
    y[row * ldy + c] = m;
    if (idx) idx[i] = (uint8_t)g;
  }
}
extern "C" int a3d_pool4_reduce_f32(a3d_ctx* ctx, const float* acc, const float* bias, float* y, int ldy, uint8_t* idx,
                                    size_t rows, unsigned flags, void* stream) {
  A3D_REQUIRE(ctx && acc && y && ldy >= 64, "pool4_reduce_f32: bad argument");
  pool4_reduce_f32_kernel<<<grid_f32(ctx, rows * 64), 256, 0, as_stream(stream)>>>(acc, bias, y, ldy, idx, rows, flags);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// dybig[row][g*64 + c] = dy[row][c] if g is the recorded arg-max and the pooled ReLU output is positive, else 0
__global__ void pool4_bwd_f32_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                     const uint8_t* __restrict__ idx, float* __restrict__ dybig, size_t rows) {
  const size_t total = rows * 256;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i >> 8;
    const int col = (int)(i & 255), g = col >> 6, c = col & 63;
    dybig[i] = (idx[row * 64 + c] == g && y[row * ldy + c] > 0.f) ? dy[row * lddy + c] : 0.f;
  }
}
extern "C" int a3d_pool4_bwd_f32(a3d_ctx* ctx, const float* dy, int lddy, const float* y, int ldy, const uint8_t* idx,
                                 float* dybig, size_t rows, void* stream) {
  A3D_REQUIRE(ctx && dy && y && idx && dybig, "pool4_bwd_f32: bad argument");
  pool4_bwd_f32_kernel<<<grid_f32(ctx, rows * 256), 256, 0, as_stream(stream)>>>(dy, lddy, y, ldy, idx, dybig, rows);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- BiasAddGrad: db[c] = sum_rows dy[row][c] (float32).  One block per 32 columns, warps stride the rows ---------------
__global__ void colsum_f32_kernel(const float* __restrict__ a, size_t rows, int C, int ld, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (size_t r = w; r < rows; r += 8) s += a[r * ld + c];
  red[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    out[c] = t;
  }
}
// long matrices: partial sums of row chunks with atomics (out zeroed first)
__global__ void colsum_f32_chunk_kernel(const float* __restrict__ a, size_t rows, int C, int ld, float* __restrict__ out,
                                        int chunk) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  const size_t r0 = (size_t)blockIdx.y * chunk;
  size_t r1 = r0 + chunk;
  if (r1 > rows) r1 = rows;
  float s = 0.f;
  if (c < C)
    for (size_t r = r0 + w; r < r1; r += 8) s += a[r * ld + c];
  __shared__ float red[8][33];
  red[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    atomicAdd(out + c, t);
  }
}
extern "C" int a3d_bias_grad_f32(a3d_ctx* ctx, const float* dy, size_t rows, int C, int ld, float* db, void* stream) {
  A3D_REQUIRE(ctx && dy && db && rows > 0 && C > 0 && ld >= C, "bias_grad_f32: bad argument");
  cudaStream_t st = as_stream(stream);
  const int gx = (C + 31) / 32;
  if (rows <= 4096) {
    colsum_f32_kernel<<<gx, 256, 0, st>>>(dy, rows, C, ld, db);
  } else {
    const int chunk = 2048;
    A3D_CHECK_CUDA(cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st));
    colsum_f32_chunk_kernel<<<dim3(gx, (unsigned)((rows + chunk - 1) / chunk)), 256, 0, st>>>(dy, rows, C, ld, db, chunk);
  }
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- dst[idx[g][e]] = src[e] (float32 copy of a3d_scatter_cast_bf16: re-embed the canonical fine/first filter) ---------
__global__ void scatter_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, int G, size_t n,
                                   float* __restrict__ dst) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const float v = src[e];
    for (int g = 0; g < G; ++g) {
      const int k = idx[(size_t)g * n + e];
      if (k >= 0) dst[k] = v;
    }
  }
}
extern "C" int a3d_scatter_f32(a3d_ctx* ctx, const float* src, const int* idx, int G, size_t n, float* dst, void* stream) {
  A3D_REQUIRE(ctx && src && idx && dst, "scatter_f32: bad argument");
  scatter_f32_kernel<<<grid_f32(ctx, n), 256, 0, as_stream(stream)>>>(src, idx, G, n, dst);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- helpers of the TF32 convolution entry points (tf32_conv.cu) -------------------------------------------------------
// w[co][t][ci] -> wd[ci][RS-1-t][co]  (dgrad of a stride-1 convolution = forward convolution with this filter)
__global__ void flip_filter_f32_kernel(const float* __restrict__ w, float* __restrict__ wd, int K, int RS, int C) {
  const size_t total = (size_t)K * RS * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % K);
    const size_t t2 = i / K;
    const int tf = (int)(t2 % RS);
    const int ci = (int)(t2 / RS);
    wd[i] = w[((size_t)co * RS + (RS - 1 - tf)) * C + ci];
  }
}
int a3d_flip_filter_f32(a3d_ctx* ctx, const float* w, float* wd, int K, int RS, int C, cudaStream_t st) {
  flip_filter_f32_kernel<<<grid_f32(ctx, (size_t)K * RS * C), 256, 0, st>>>(w, wd, K, RS, C);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// dx[n,ih,iw,c] = sum over the taps (r,s) that reach (ih,iw) of col[(n,p,q)][(r,s,c)]   (strided dgrad, second half)
__global__ void col2im_f32_kernel(const float* __restrict__ col, float* __restrict__ dx, int N, int H, int W, int C, int R,
                                  int S, int sh, int sw, int pt, int pl, int P, int Q) {
  const size_t J = (size_t)R * S * C;
  const size_t total = (size_t)N * H * W * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n, ih, iw, c;
    split_nhwc(i, H, W, C, n, ih, iw, c);
    float acc = 0.f;
    for (int r = 0; r < R; ++r) {
      const int ph = ih + pt - r;
      if (ph < 0 || ph % sh) continue;
      const int p = ph / sh;
      if (p >= P) continue;
      for (int s = 0; s < S; ++s) {
        const int qw = iw + pl - s;
        if (qw < 0 || qw % sw) continue;
        const int q = qw / sw;
        if (q >= Q) continue;
        acc += __ldg(col + (((size_t)n * P + p) * Q + q) * J + (size_t)(r * S + s) * C + c);
      }
    }
    dx[i] = acc;
  }
}
int a3d_col2im_f32(a3d_ctx* ctx, const float* col, float* dx, const a3d_conv_desc* d, cudaStream_t st) {
  col2im_f32_kernel<<<grid_f32(ctx, (size_t)d->N * d->H * d->W * d->C), 256, 0, st>>>(
      col, dx, d->N, d->H, d->W, d->C, d->R, d->S, d->stride_h, d->stride_w, d->pad_t, d->pad_l, d->P, d->Q);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- single-filter convolution in exact float32 (MSDN fine/third, src/models.py:250: 5x5x64 -> 1, SAME) -----------------
// A 128-row tensor-core tile would be 127/128 padding and its dY / filter "matrices" have a 4-byte pitch (no TMA): one
// warp per pixel, lanes split the channels (coalesced 128 B per tap), warp-shuffle reductions.  HBM/L2-bound.
__global__ void conv_k1_fwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                       float* __restrict__ y, int N, int H, int W, int C, int R, int S, int sh, int sw, int pt,
                                       int pl, int P, int Q, int ldy, unsigned flags) {
  extern __shared__ float wsm[];
  for (int i = threadIdx.x; i < R * S * C; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long total = (long long)N * P * Q;
  for (long long m = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); m < total; m += (long long)gridDim.x * wpb) {
    const int q = (int)(m % Q);
    const long long t = m / Q;
    const int p = (int)(t % P), n = (int)(t / P);
    float acc = 0.f;
    for (int r = 0; r < R; ++r) {
      const int ih = p * sh - pt + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < S; ++s) {
        const int iw = q * sw - pl + s;
        if (iw < 0 || iw >= W) continue;
        const float* px = x + (((size_t)n * H + ih) * W + iw) * C;
        const float* wf = wsm + (r * S + s) * C;
        for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(px + c), wf[c], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (bias) acc += bias[0];
      if (flags & A3D_EPI_RELU) acc = fmaxf(acc, 0.f);
      y[(size_t)m * ldy] = acc;
    }
  }
}

// dx[n,ih,iw,c] = relu'(src) * sum_taps dy[n,p,q] w[r][s][c]   with p*sh - pt + r = ih, q*sw - pl + s = iw
__global__ void conv_k1_dgrad_f32_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                         const float* __restrict__ relu_src, int N, int H, int W, int C, int R, int S, int sh,
                                         int sw, int pt, int pl, int P, int Q, int ldy) {
  extern __shared__ float wsm[];
  for (int i = threadIdx.x; i < R * S * C; i += blockDim.x) wsm[i] = w[i];
  __syncthreads();
  const size_t total = (size_t)N * H * W * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int n, ih, iw, c;
    split_nhwc(i, H, W, C, n, ih, iw, c);
    float acc = 0.f;
    for (int r = 0; r < R; ++r) {
      const int ph = ih + pt - r;
      if (ph < 0 || ph % sh) continue;
      const int p = ph / sh;
      if (p >= P) continue;
      for (int s = 0; s < S; ++s) {
        const int qw = iw + pl - s;
        if (qw < 0 || qw % sw) continue;
        const int q = qw / sw;
        if (q >= Q) continue;
        acc = fmaf(__ldg(dy + (((size_t)n * P + p) * Q + q) * ldy), wsm[(r * S + s) * C + c], acc);
      }
    }
    if (relu_src && !(relu_src[i] > 0.f)) acc = 0.f;
    dx[i] = acc;
  }
}

// dw[r][s][c] = sum_pixels dy[pix] x[pix @ tap (r,s)][c] ; db = sum dy.  One block per chunk of output pixels, partial
// sums per (tap, channel) in registers of the (tap-strided) threads, float atomics into the zeroed dw.
__global__ void conv_k1_wgrad_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                         float* __restrict__ db, int N, int H, int W, int C, int R, int S, int sh, int sw,
                                         int pt, int pl, int P, int Q, int ldy, int chunk) {
  const long long total = (long long)N * P * Q;
  const long long m0 = (long long)blockIdx.x * chunk;
  long long m1 = m0 + chunk;
  if (m1 > total) m1 = total;
  const int J = R * S * C;
  float bsum = 0.f;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    const int tap = j / C, c = j - tap * C;
    const int r = tap / S, s = tap - r * S;
    float acc = 0.f;
    for (long long m = m0; m < m1; ++m) {
      const int q = (int)(m % Q);
      const long long t = m / Q;
      const int p = (int)(t % P), n = (int)(t / P);
      const int ih = p * sh - pt + r, iw = q * sw - pl + s;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      acc = fmaf(__ldg(dy + (size_t)m * ldy), __ldg(x + (((size_t)n * H + ih) * W + iw) * C + c), acc);
    }
    atomicAdd(dw + j, acc);
  }
  if (db && threadIdx.x == 0) {
    for (long long m = m0; m < m1; ++m) bsum += dy[(size_t)m * ldy];
    atomicAdd(db, bsum);
  }
}

static int k1_check(a3d_ctx* ctx, const a3d_conv_desc* d) {
  A3D_REQUIRE(ctx && d && d->K == 1 && (size_t)d->R * d->S * d->C * 4 <= 48 * 1024, "conv_k1 f32: needs K == 1 and a filter of <= 48 KB");
  return 0;
}
extern "C" int a3d_conv_k1_fwd_f32(a3d_ctx* ctx, const a3d_conv_desc* d, const float* x, const float* w, const float* bias,
                                   float* y, unsigned flags, void* stream) {
  int rc = k1_check(ctx, d);
  if (rc) return rc;
  const long long pixels = (long long)d->N * d->P * d->Q;
  int grid = (int)((pixels + 7) / 8);
  if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
  conv_k1_fwd_f32_kernel<<<grid, 256, (size_t)d->R * d->S * d->C * 4, as_stream(stream)>>>(
      x, w, bias, y, d->N, d->H, d->W, d->C, d->R, d->S, d->stride_h, d->stride_w, d->pad_t, d->pad_l, d->P, d->Q, d->ldy, flags);
  A3D_LAUNCH_OK(ctx);
  return 0;
}
extern "C" int a3d_conv_k1_dgrad_f32(a3d_ctx* ctx, const a3d_conv_desc* d, const float* dy, const float* w, float* dx,
                                     const float* relu_src, void* stream) {
  int rc = k1_check(ctx, d);
  if (rc) return rc;
  conv_k1_dgrad_f32_kernel<<<grid_f32(ctx, (size_t)d->N * d->H * d->W * d->C), 256, (size_t)d->R * d->S * d->C * 4,
                             as_stream(stream)>>>(dy, w, dx, relu_src, d->N, d->H, d->W, d->C, d->R, d->S, d->stride_h,
                                                  d->stride_w, d->pad_t, d->pad_l, d->P, d->Q, d->ldy);
  A3D_LAUNCH_OK(ctx);
  return 0;
}
extern "C" int a3d_conv_k1_wgrad_f32(a3d_ctx* ctx, const a3d_conv_desc* d, const float* x, const float* dy, float* dw,
                                     float* db, void* stream) {
  int rc = k1_check(ctx, d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const long long pixels = (long long)d->N * d->P * d->Q;
  const int J = d->R * d->S * d->C;
  A3D_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)J * sizeof(float), st));
  if (db) A3D_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float), st));
  int blocks = ctx->sm_count * 4;
  int chunk = (int)((pixels + blocks - 1) / blocks);
  if (chunk < 1) chunk = 1;
  blocks = (int)((pixels + chunk - 1) / chunk);
  conv_k1_wgrad_f32_kernel<<<blocks, 256, 0, st>>>(x, dy, dw, db, d->N, d->H, d->W, d->C, d->R, d->S, d->stride_h, d->stride_w,
                                                   d->pad_t, d->pad_l, d->P, d->Q, d->ldy, chunk);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- 3xTF32 operand split ---------------------------------------------------------------------------------------
// hi = x rounded to TF32 (10 mantissa bits, cvt.rna), lo = x - hi (exact in float32; |lo| <= 2^-11 |x|).  The tensor core
// then sees hi unchanged and lo to 10 more bits, so  hi_a.hi_b + hi_a.lo_b + lo_a.hi_b  carries ~21 mantissa bits of a.b.
// Rows may be pitched (ld floats between rows, `cols` used per row); hi / lo are written densely ([rows][cols]).
__global__ void split_tf32_kernel(const float* __restrict__ x, size_t rows, int cols, long long ld, float* __restrict__ hi,
                                  float* __restrict__ lo) {
  const size_t n = rows * (size_t)cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float v = x[r * ld + c];
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float h = __uint_as_float(hb);
    hi[i] = h;
    lo[i] = v - h;
  }
}

extern "C" int a3d_split_tf32(a3d_ctx* ctx, const float* x, size_t rows, int cols, long long ld, float* hi, float* lo,
                              void* stream) {
  A3D_REQUIRE(ctx && x && hi && lo && cols > 0 && ld >= cols, "split_tf32: bad argument");
  if (!rows) return 0;
  const size_t n = rows * (size_t)cols;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
  split_tf32_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, rows, cols, ld, hi, lo);
  A3D_LAUNCH_OK(ctx);
  return 0;
}
