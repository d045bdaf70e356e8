// crf.cu -- DCNF structured part (src/models.py:95-177): fused pairwise features, tile means,
// patch gather and the batched closed-form CRF (Cholesky in shared memory, one CTA per graph).
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int CRF_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

// A (n x n, row-major, pitch n+1 to dodge bank conflicts) lives in dynamic smem.
__global__ void __launch_bounds__(CRF_THREADS)
crf_kernel(const float* __restrict__ z_, const float* __restrict__ y_, const float* __restrict__ r_,
           const int32_t* __restrict__ pl, const int32_t* __restrict__ pr, int n, int n_pairs, float grad_scale,
           int naive, float* __restrict__ ystar_, float* __restrict__ nll_, float* __restrict__ logdet_, float* __restrict__ dz_,
           float* __restrict__ dr_, int32_t* __restrict__ status_) {
  extern __shared__ float sm[];
  const int ld = n + 1;
  float* A = sm;                       // n*ld
  float* z = A + (size_t)n * ld;       // n
  float* y = z + n;                    // n
  float* w = y + n;                    // n   (L w = z)
  float* ys = w + n;                   // n   (L^T ys = w)
  float* red = ys + n;                 // 32
  float* tmp = red + 32;               // (warps) * n scratch for dr solves
  __shared__ int s_status;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* r = r_ + (size_t)b * n_pairs;

  for (int i = tid; i < n * ld; i += nt) A[i] = 0.f;
  for (int i = tid; i < n; i += nt) { z[i] = z_[(size_t)b * n + i]; y[i] = y_[(size_t)b * n + i]; }
  if (tid == 0) s_status = 0;
  __syncthreads();
  // R[l][r] = R[r][l] = r_k  (scatter *update* semantics of src/models.py:138-141: last write wins)
  if (tid == 0)
    for (int k = 0; k < n_pairs; ++k) {
      int l = pl[k], q = pr[k];
      A[l * ld + q] = r[k];
      A[q * ld + l] = r[k];
    }
  __syncthreads();
  // A = I + diag(R 1) - R
  for (int i = tid; i < n; i += nt) {
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += A[i * ld + j];
    w[i] = s;                          // row sums (A still holds R)
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += nt) {
    int i = idx / n, j = idx - i * n;
    float v = -A[i * ld + j];
    if (i == j) v += 1.f + w[i];
    A[i * ld + j] = v;
  }
  __syncthreads();
  // energy = y^T A y - 2 z^T y + z^T z
  float e = 0.f, zz = 0.f;
  for (int i = tid; i < n; i += nt) {
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += A[i * ld + j] * y[j];
    e += y[i] * s - 2.f * z[i] * y[i] + z[i] * z[i];
    zz += z[i] * z[i];
  }
  const float energy = block_sum(e, red);
  const float ztz = block_sum(zz, red);

  // in-place right-looking Cholesky (lower triangle): A = L L^T
  for (int k = 0; k < n; ++k) {
    if (tid == 0) {
      float d = A[k * ld + k];
      if (!(d > 0.f)) { if (s_status == 0) s_status = k + 1; d = 1.f; }
      A[k * ld + k] = sqrtf(d);
    }
    __syncthreads();
    const float dk = A[k * ld + k];
    for (int i = k + 1 + tid; i < n; i += nt) A[i * ld + k] /= dk;
    __syncthreads();
    const int rem = n - k - 1;
    for (int idx = tid; idx < rem * rem; idx += nt) {
      int i = k + 1 + idx / rem, j = k + 1 + idx % rem;
      if (j <= i) A[i * ld + j] -= A[i * ld + k] * A[j * ld + k];
    }
    __syncthreads();
  }
  float ldsum = 0.f;
  for (int i = tid; i < n; i += nt) ldsum += logf(A[i * ld + i]);
  const float logdet = 2.f * block_sum(ldsum, red);

  // forward / backward substitution by warp 0 (dot products via shuffles)
  if (tid < 32) {
    for (int i = 0; i < n; ++i) {
      float s = 0.f;
      for (int j = tid; j < i; j += 32) s += A[i * ld + j] * w[j];
      s = warp_sum(s);
      if (tid == 0) w[i] = (z[i] - s) / A[i * ld + i];
      __syncwarp();
    }
    for (int i = n - 1; i >= 0; --i) {
      float s = 0.f;
      for (int j = i + 1 + tid; j < n; j += 32) s += A[j * ld + i] * ys[j];
      s = warp_sum(s);
      if (tid == 0) ys[i] = (w[i] - s) / A[i * ld + i];
      __syncwarp();
    }
  }
  __syncthreads();
  float q = 0.f;
  for (int i = tid; i < n; i += nt) q += w[i] * w[i];
  const float quad = block_sum(q, red);
  const bool ok = (s_status == 0);
  float nll = energy + 0.5f * n * logf(CUDART_PI_F) - 0.5f * logdet + quad - ztz;
  if (naive) {
    // src/models.py:163-171 evaluated literally in f32 (eps = 1e-7 everywhere)
    const float eps = 1e-7f;
    float zs = 0.f;
    for (int i = tid; i < n; i += nt) zs += z[i];
    const float zsum = block_sum(zs, red);
    float fac = powf(CUDART_PI_F, 0.5f * n) / (sqrtf(expf(logdet)) + eps);
    float Z = fac * expf(quad + eps * zsum * zsum - ztz) + eps;
    float u = expf(-energy) / Z;
    nll = -logf(u + eps);
    grad_scale *= u / (u + eps);
  }
  if (tid == 0) {
    status_[b] = s_status;
    nll_[b] = ok ? nll : 0.f;
    if (logdet_) logdet_[b] = ok ? logdet : 0.f;
  }
  for (int i = tid; i < n; i += nt) {
    if (ystar_) ystar_[(size_t)b * n + i] = ok ? ys[i] : 0.f;
    if (dz_) dz_[(size_t)b * n + i] = ok ? grad_scale * 2.f * (ys[i] - y[i]) : 0.f;
  }
  if (dr_) {
    // d nll / d r_k = (y_l-y_r)^2 - (y*_l-y*_r)^2 - 0.5 * || L^-1 (e_l - e_r) ||^2
    const int wid = tid >> 5, lane = tid & 31, nw = nt >> 5;
    float* t = tmp + (size_t)wid * n;
    for (int k = wid; k < n_pairs; k += nw) {
      int l = pl[k], rr = pr[k];
      float nrm = 0.f;
      for (int i = 0; i < n; ++i) {
        float s = 0.f;
        for (int j = lane; j < i; j += 32) s += A[i * ld + j] * t[j];
        s = warp_sum(s);
        float rhs = (i == l ? 1.f : 0.f) - (i == rr ? 1.f : 0.f);
        float ti = (rhs - s) / A[i * ld + i];
        if (lane == 0) t[i] = ti;
        nrm += ti * ti;
        __syncwarp();
      }
      if (lane == 0) {
        float dy = y[l] - y[rr], ds = ys[l] - ys[rr];
        dr_[(size_t)b * n_pairs + k] = ok ? grad_scale * (dy * dy - ds * ds - 0.5f * nrm) : 0.f;
      }
    }
  }
}

// ---- pairwise features ---------------------------------------------------------------------
constexpr int TILE = 40, TPIX = TILE * TILE, NBINS = 256, FEAT = TPIX + NBINS;

// one CTA per (image, tile): channel-mean vector [1600] and 256-bin colour histogram
__global__ void tile_features_kernel(const float* __restrict__ img, int H, int W, int cols, int n_tiles,
                                     float* __restrict__ feat) {
  __shared__ unsigned int hist[NBINS];
  const int b = blockIdx.x / n_tiles, t = blockIdx.x % n_tiles;
  const int tr = t / cols, tcol = t % cols;
  for (int i = threadIdx.x; i < NBINS; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  float* f = feat + (size_t)blockIdx.x * FEAT;
  for (int i = threadIdx.x; i < TPIX; i += blockDim.x) {
    int py = tr * TILE + i / TILE, px = tcol * TILE + i % TILE;
    float r = 0.f, g = 0.f, bl = 0.f;
    if (py < H && px < W) {       // SAME padding of extract_image_patches: zeros outside
      const float* p = img + (((size_t)b * H + py) * W + px) * 3;
      r = p[0]; g = p[1]; bl = p[2];
    }
    f[i] = (r + g + bl) / 3.f;
    // src/models.py:97-99: v = R*2^24 + G*2^16 + B*2^8 ; bin = clip(floor(256 * v / 2^24), 0, 255)
    float v = r * 16777216.f + g * 65536.f + bl * 256.f;
    float scaled = v / 16777216.f;
    int bin = (int)floorf(scaled * (float)NBINS);
    bin = max(0, min(NBINS - 1, bin));
    atomicAdd(&hist[bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NBINS; i += blockDim.x) f[TPIX + i] = (float)hist[i];
}

// one warp per (image, pair)
__global__ void pair_similarity_kernel(const float* __restrict__ feat, const int32_t* __restrict__ pl,
                                       const int32_t* __restrict__ pr, int n_tiles, int n_pairs, int total, float gamma,
                                       float* __restrict__ sims) {
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= total) return;
  int b = gw / n_pairs, k = gw % n_pairs;
  const float* fl = feat + ((size_t)b * n_tiles + pl[k]) * FEAT;
  const float* fr = feat + ((size_t)b * n_tiles + pr[k]) * FEAT;
  float s0 = 0.f, s1 = 0.f;
  for (int i = lane; i < TPIX; i += 32) { float d = fl[i] - fr[i]; s0 += d * d; }
  for (int i = lane; i < NBINS; i += 32) { float d = fl[TPIX + i] - fr[TPIX + i]; s1 += d * d; }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    sims[(size_t)gw * 2 + 0] = expf(-gamma * sqrtf(s0));
    sims[(size_t)gw * 2 + 1] = expf(-gamma * sqrtf(s1));
  }
}

__global__ void tile_means_kernel(const float* __restrict__ depth, int H, int W, int cols, int n_tiles,
                                  float* __restrict__ y) {
  __shared__ float red[32];
  const int b = blockIdx.x / n_tiles, t = blockIdx.x % n_tiles;
  const int tr = t / cols, tcol = t % cols;
  float s = 0.f;
  for (int i = threadIdx.x; i < TPIX; i += blockDim.x) {
    int py = tr * TILE + i / TILE, px = tcol * TILE + i % TILE;
    if (py < H && px < W) s += depth[((size_t)b * H + py) * W + px];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) y[blockIdx.x] = s / (float)TPIX;
}

// 100x100 patches, stride 40, SAME (30 px zero border) -> bf16 NHWC with dstC channels
__global__ void extract_patches_kernel(const float* __restrict__ img, int B, int H, int W, int rows, int cols,
                                       uint16_t* __restrict__ out, int dstC) {
  const int PS = 100, PAD = 30;
  size_t total = (size_t)B * rows * cols * PS * PS;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int px = (int)(i % PS);
    size_t t = i / PS;
    int py = (int)(t % PS);
    t /= PS;
    int pc = (int)(t % cols);
    t /= cols;
    int prow = (int)(t % rows);
    int b = (int)(t / rows);
    int iy = prow * TILE - PAD + py, ix = pc * TILE - PAD + px;
    float v[3] = {0.f, 0.f, 0.f};
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float* p = img + (((size_t)b * H + iy) * W + ix) * 3;
      v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    }
    uint16_t* o = out + i * dstC;
    for (int c = 0; c < dstC; ++c) o[c] = c < 3 ? f32_to_bf16_bits(v[c]) : (uint16_t)0;
  }
}

// The same patches after space-to-depth(2): 50 x 50 cells of 16 channels, channel (2a + b) * 3 + c = patch pixel
// (2Y + a, 2X + b, c), channels 12..15 zero.  fold = 1: out [NP][50][50][16] (one 32-byte cell per position).
// fold = 4: out [NP][50][50][64], position X holds cells X .. X+3 (zeros beyond the row) -- the materialised form of the
// overlapped view the first DCNF layer reads (a3d_conv_desc::pix_pitch = 16 over the fold = 1 tensor is the same data).
__global__ void extract_patches_s2d_kernel(const float* __restrict__ img, int B, int H, int W, int rows, int cols,
                                           uint16_t* __restrict__ out, int fold) {
  const int CS = 50, PAD = 30;
  size_t total = (size_t)B * rows * cols * CS * CS * fold;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % fold);
    size_t t = i / fold;
    const int X = (int)(t % CS);
    t /= CS;
    const int Y = (int)(t % CS);
    t /= CS;
    const int pc = (int)(t % cols);
    t /= cols;
    const int prow = (int)(t % rows);
    const int b = (int)(t / rows);
    uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int cx = X + k;
    if (cx < CS) {
      float f[12];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int iy = prow * TILE - PAD + 2 * Y + a, ix = pc * TILE - PAD + 2 * cx + bb;
          const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
          const float* p = img + (((size_t)b * H + (ok ? iy : 0)) * W + (ok ? ix : 0)) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) f[(a * 2 + bb) * 3 + c] = ok ? p[c] : 0.f;
        }
#pragma unroll
      for (int j = 0; j < 6; ++j) v[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = make_uint4(v[0], v[1], v[2], v[3]);
    o[1] = make_uint4(v[4], v[5], v[6], v[7]);
  }
}

// ---- fully convolutional unary network (DCNF): no patches at all ---------------------------------------------------------
// The reference runs its CNN on 48 overlapping 100x100 patches per image (src/models.py:50-83).  Patches are plain windows
// (stride 40, zero border 30) and every layer is a VALID convolution or an even-aligned 2x2 pool, so layer L of patch
// (prow, pcol) IS a window of layer L of the zero-padded whole image: the network is evaluated ONCE per image (a third of
// the FLOPs) and only the 7x7x256 inputs of the first dense layer are gathered per patch (windows of stride 5).
// image_cells_s2d: zero-padded image after space-to-depth(2): cells bf16 [B][(H+2pad)/2][(W+2pad)/2][16], channel
// (2a+b)*3+c = padded pixel (2Y+a, 2X+b, c), channels 12..15 zero.
__global__ void image_cells_s2d_kernel(const float* __restrict__ img, int B, int H, int W, int pad, int CH, int CW,
                                       uint16_t* __restrict__ out) {
  size_t total = (size_t)B * CH * CW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % CW);
    size_t t = i / CW;
    const int Y = (int)(t % CH);
    const int b = (int)(t / CH);
    float f[12];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int iy = 2 * Y + a - pad, ix = 2 * X + bb - pad;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const float* p = img + (((size_t)b * H + (ok ? iy : 0)) * W + (ok ? ix : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) f[(a * 2 + bb) * 3 + c] = ok ? p[c] : 0.f;
      }
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    o[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), 0u, 0u);
  }
}

// out[(b, pr, pc)][i][j][:] = src[b][pr*stride + i][pc*stride + j][:]   (16-byte pieces of 8 channels)
__global__ void window_gather_kernel(const uint4* __restrict__ src, int B, int Hs, int Ws, int C8, int rows, int cols, int win,
                                     int stride, uint4* __restrict__ out) {
  size_t total = (size_t)B * rows * cols * win * win * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t t = i / C8;
    const int j = (int)(t % win); t /= win;
    const int ii = (int)(t % win); t /= win;
    const int pc = (int)(t % cols); t /= cols;
    const int pr = (int)(t % rows);
    const int b = (int)(t / rows);
    out[i] = __ldg(src + (((size_t)b * Hs + pr * stride + ii) * Ws + pc * stride + j) * C8 + c);
  }
}

// transpose of the gather: g_src[b][h][w][:] = sum over the windows that contain (h, w) of g_out[(b,pr,pc)][h-pr*stride][w-pc*stride][:]
// (float32 sum, one bf16 rounding; gather form, no atomics)
__global__ void window_scatter_sum_kernel(const uint4* __restrict__ g_out, int B, int Hs, int Ws, int C8, int rows, int cols,
                                          int win, int stride, uint4* __restrict__ g_src) {
  size_t total = (size_t)B * Hs * Ws * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    size_t t = i / C8;
    const int w = (int)(t % Ws); t /= Ws;
    const int h = (int)(t % Hs);
    const int b = (int)(t / Hs);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int pr = 0; pr < rows; ++pr) {
      const int ii = h - pr * stride;
      if (ii < 0 || ii >= win) continue;
      for (int pc = 0; pc < cols; ++pc) {
        const int j = w - pc * stride;
        if (j < 0 || j >= win) continue;
        const uint4 v = __ldg(g_out + ((((size_t)(b * rows + pr) * cols + pc) * win + ii) * win + j) * C8 + c);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[2 * k] += __uint_as_float(u[k] << 16);
          acc[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
        }
      }
    }
    g_src[i] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                          pack_bf16x2(acc[6], acc[7]));
  }
}

}  // namespace

extern "C" int a3d_image_cells_s2d(a3d_ctx* ctx, const float* images, int B, int H, int W, int pad, uint16_t* cells,
                                   void* stream) {
  A3D_REQUIRE(ctx && images && cells && B > 0 && pad >= 0 && (H + 2 * pad) % 2 == 0 && (W + 2 * pad) % 2 == 0,
              "image_cells_s2d: bad argument (padded size must be even)");
  A3D_REQUIRE((reinterpret_cast<uintptr_t>(cells) & 15) == 0, "image_cells_s2d: output must be 16-byte aligned");
  const int CH = (H + 2 * pad) / 2, CW = (W + 2 * pad) / 2;
  size_t total = (size_t)B * CH * CW;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
  image_cells_s2d_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(images, B, H, W, pad, CH, CW, cells);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int a3d_window_gather(a3d_ctx* ctx, const uint16_t* src, int B, int Hs, int Ws, int C, int rows, int cols, int win,
                                 int stride, uint16_t* out, void* stream) {
  A3D_REQUIRE(ctx && src && out && B > 0 && C % 8 == 0 && rows > 0 && cols > 0 && win > 0 && stride > 0 &&
                  (rows - 1) * stride + win <= Hs && (cols - 1) * stride + win <= Ws,
              "window_gather: bad argument (C %% 8 == 0, windows inside the source)");
  size_t total = (size_t)B * rows * cols * win * win * (C / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
  window_gather_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(src), B, Hs, Ws, C / 8, rows,
                                                                   cols, win, stride, reinterpret_cast<uint4*>(out));
  A3D_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int a3d_window_scatter_sum(a3d_ctx* ctx, const uint16_t* g_out, int B, int Hs, int Ws, int C, int rows, int cols,
                                      int win, int stride, uint16_t* g_src, void* stream) {
  A3D_REQUIRE(ctx && g_out && g_src && B > 0 && C % 8 == 0 && rows > 0 && cols > 0 && win > 0 && stride > 0 &&
                  (rows - 1) * stride + win <= Hs && (cols - 1) * stride + win <= Ws,
              "window_scatter_sum: bad argument (C %% 8 == 0, windows inside the source)");
  size_t total = (size_t)B * Hs * Ws * (C / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
  window_scatter_sum_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(g_out), B, Hs, Ws, C / 8,
                                                                        rows, cols, win, stride,
                                                                        reinterpret_cast<uint4*>(g_src));
  A3D_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int a3d_extract_patches_s2d(a3d_ctx* ctx, const float* images, int B, int H, int W, uint16_t* cells, int fold,
                                       void* stream) {
  A3D_REQUIRE(ctx && images && cells && (fold == 1 || fold == 4), "extract_patches_s2d: bad argument (fold is 1 or 4)");
  A3D_REQUIRE((reinterpret_cast<uintptr_t>(cells) & 15) == 0, "extract_patches_s2d: output must be 16-byte aligned");
  int rows = (H + TILE - 1) / TILE, cols = (W + TILE - 1) / TILE;
  size_t total = (size_t)B * rows * cols * 50 * 50 * fold;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
  extract_patches_s2d_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(images, B, H, W, rows, cols, cells, fold);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int a3d_crf_fwd_bwd(a3d_ctx* ctx, const float* z, const float* y, const float* r, const int32_t* pl,
                               const int32_t* pr, int B, int n, int n_pairs, float grad_scale, int naive, float* ystar,
                               float* nll, float* logdet, float* dz, float* dr, int32_t* status, void* stream) {
  A3D_REQUIRE(ctx && z && y && r && pl && pr && nll && status, "crf: null argument");
  A3D_REQUIRE(B > 0 && n > 0 && n <= 192 && n_pairs >= 0, "crf: n must be in 1..192");
  size_t smem = ((size_t)n * (n + 1) + 4 * (size_t)n + 32 + (size_t)(CRF_THREADS / 32) * n) * sizeof(float);
  if (smem > 48 * 1024)
    A3D_CHECK_CUDA(cudaFuncSetAttribute(crf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  crf_kernel<<<B, CRF_THREADS, smem, as_stream(stream)>>>(z, y, r, pl, pr, n, n_pairs, grad_scale, naive, ystar, nll,
                                                         logdet, dz, dr, status);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

extern "C" size_t a3d_pairwise_ws_bytes(int B, int H, int W) {
  int rows = (H + TILE - 1) / TILE, cols = (W + TILE - 1) / TILE;
  return (size_t)B * rows * cols * FEAT * sizeof(float);
}

extern "C" int a3d_pairwise_features(a3d_ctx* ctx, const float* images, int B, int H, int W, const int32_t* pl,
                                     const int32_t* pr, int n_pairs, float gamma, float* tile_feat_ws, float* sims,
                                     void* stream) {
  A3D_REQUIRE(ctx && images && pl && pr && tile_feat_ws && sims, "pairwise_features: null argument");
  int rows = (H + TILE - 1) / TILE, cols = (W + TILE - 1) / TILE, n_tiles = rows * cols;
  tile_features_kernel<<<B * n_tiles, 256, 0, as_stream(stream)>>>(images, H, W, cols, n_tiles, tile_feat_ws);
  A3D_LAUNCH_OK(ctx);
  int total = B * n_pairs;
  pair_similarity_kernel<<<ceil_div((long long)total * 32, 256), 256, 0, as_stream(stream)>>>(
      tile_feat_ws, pl, pr, n_tiles, n_pairs, total, gamma, sims);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int a3d_tile_means(a3d_ctx* ctx, const float* depth, int B, int H, int W, float* y, void* stream) {
  A3D_REQUIRE(ctx && depth && y, "tile_means: null argument");
  int rows = (H + TILE - 1) / TILE, cols = (W + TILE - 1) / TILE, n_tiles = rows * cols;
  tile_means_kernel<<<B * n_tiles, 256, 0, as_stream(stream)>>>(depth, H, W, cols, n_tiles, y);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int a3d_extract_patches(a3d_ctx* ctx, const float* images, int B, int H, int W, uint16_t* patches, int dstC,
                                   void* stream) {
  A3D_REQUIRE(ctx && images && patches && dstC >= 3, "extract_patches: bad argument");
  int rows = (H + TILE - 1) / TILE, cols = (W + TILE - 1) / TILE;
  size_t total = (size_t)B * rows * cols * 100 * 100;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
  extract_patches_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(images, B, H, W, rows, cols, patches, dstC);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

__global__ void pairwise_dense_kernel(const float* __restrict__ sims, const float* __restrict__ w, const float* __restrict__ b,
                                      float* __restrict__ r, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    r[i] = sims[2 * i] * w[0] + sims[2 * i + 1] * w[1] + b[0];
}
extern "C" int a3d_pairwise_dense(a3d_ctx* ctx, const float* sims, const float* w2, const float* b1, float* r, size_t n,
                                  void* stream) {
  A3D_REQUIRE(ctx && sims && w2 && b1 && r, "pairwise_dense: null argument");
  pairwise_dense_kernel<<<ceil_div((long long)n, 256), 256, 0, as_stream(stream)>>>(sims, w2, b1, r, n);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// ---- beyond the reference: a non-negative pairwise potential (A = I + D - R is SPD for every r >= 0;
// the reference's unconstrained linear layer, src/models.py:92-93, can leave the SPD cone) and the gradient INTO the
// pairwise layer (TF 1.3 blocks it at ScatterNdUpdate; the CRF kernel's `dr` output is the gradient w.r.t. r).
//   fwd: r = act(sims . w + b), act = max(., 0) with A3D_EPI_RELU, identity otherwise
//   bwd: dw[j] = sum_i dr[i] act'(r[i]) sims[i][j], db = sum_i dr[i] act'(r[i])        (one block, n = B * n_pairs is small)
__global__ void pairwise_dense_act_kernel(const float* __restrict__ sims, const float* __restrict__ w, const float* __restrict__ b,
                                          float* __restrict__ r, size_t n, unsigned flags) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = sims[2 * i] * w[0] + sims[2 * i + 1] * w[1] + b[0];
    if (flags & A3D_EPI_RELU) v = fmaxf(v, 0.f);
    r[i] = v;
  }
}
extern "C" int a3d_pairwise_dense_act(a3d_ctx* ctx, const float* sims, const float* w2, const float* b1, float* r, size_t n,
                                      unsigned flags, void* stream) {
  A3D_REQUIRE(ctx && sims && w2 && b1 && r, "pairwise_dense_act: null argument");
  pairwise_dense_act_kernel<<<ceil_div((long long)n, 256), 256, 0, as_stream(stream)>>>(sims, w2, b1, r, n, flags);
  A3D_LAUNCH_OK(ctx);
  return 0;
}
__global__ void pairwise_dense_bwd_kernel(const float* __restrict__ sims, const float* __restrict__ r, const float* __restrict__ dr,
                                          float* __restrict__ dw2, float* __restrict__ db1, int n, unsigned flags) {
  __shared__ float red[32];
  float s0 = 0.f, s1 = 0.f, sb = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float g = dr[i];
    if ((flags & A3D_EPI_RELU) && !(r[i] > 0.f)) g = 0.f;
    s0 += g * sims[2 * i];
    s1 += g * sims[2 * i + 1];
    sb += g;
  }
  s0 = block_sum(s0, red); __syncthreads();
  s1 = block_sum(s1, red); __syncthreads();
  sb = block_sum(sb, red);
  if (threadIdx.x == 0) { dw2[0] = s0; dw2[1] = s1; db1[0] = sb; }
}
extern "C" int a3d_pairwise_dense_bwd(a3d_ctx* ctx, const float* sims, const float* r, const float* dr, float* dw2, float* db1,
                                      size_t n, unsigned flags, void* stream) {
  A3D_REQUIRE(ctx && sims && r && dr && dw2 && db1 && n > 0 && n < (1u << 30), "pairwise_dense_bwd: bad argument");
  pairwise_dense_bwd_kernel<<<1, 256, 0, as_stream(stream)>>>(sims, r, dr, dw2, db1, (int)n, flags);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

__global__ void mean_f32_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s / (float)n;
}
extern "C" int a3d_mean_f32(a3d_ctx* ctx, const float* v, int n, float* out, void* stream) {
  A3D_REQUIRE(ctx && v && out && n > 0, "mean: bad argument");
  mean_f32_kernel<<<1, 256, 0, as_stream(stream)>>>(v, n, out);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

__global__ void scale_cast_bf16_kernel(const float* __restrict__ s, uint16_t* __restrict__ d, size_t n, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    d[i] = f32_to_bf16_bits(s[i] * scale);
}
extern "C" int a3d_scale_cast_bf16(a3d_ctx* ctx, const float* src, uint16_t* dst, size_t n, float scale, void* stream) {
  A3D_REQUIRE(ctx && src && dst, "scale_cast: null argument");
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
  scale_cast_bf16_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src, dst, n, scale);
  A3D_LAUNCH_OK(ctx);
  return 0;
}
