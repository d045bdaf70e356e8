// simt_gemm.cu -- CUDA-core implicit-GEMM convolution (fwd / dgrad / wgrad) and dense kernels.
// This is the shape-agnostic path: it takes every stride / padding / channel count, and is what
// A3D_IMPL_SIMT selects.  The tcgen05 kernels in tc_gemm.cu are validated against it on-device.
// bf16 operands, fp32 accumulation, one 64x64 output tile per CTA, 4x4 outputs per thread.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NT = 256;

// Problem functors: C[i][j] = sum_k A(i,k) * B(j,k)
struct ConvFwdProb {
  const uint16_t* x; const uint16_t* w; const float* bias; void* y;
  int N, H, W, C, K, R, S, sh, sw, pt, pl, P, Q, ldy, y_f32; unsigned flags;
  int dil, pp;                 // horizontal tap spacing, pixel pitch in elements (a3d_conv_desc dil_w / pix_pitch)
  __device__ int dimM() const { return N * P * Q; }
  __device__ int dimN() const { return K; }
  __device__ int dimK() const { return R * S * C; }
  __device__ float a(int m, int k) const {
    int c = k % C; int t = k / C; int s = t % S; int r = t / S;
    int q = m % Q; int t2 = m / Q; int p = t2 % P; int n = t2 / P;
    int ih = p * sh - pt + r, iw = q * sw - pl + s * dil;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0.f;
    return bf16_bits_to_f32(x[(((size_t)n * H + ih) * W + iw) * pp + c]);
  }
  __device__ float b(int j, int k) const { return bf16_bits_to_f32(w[(size_t)j * (R * S * C) + k]); }
  __device__ void store(int m, int j, float acc) const {
    if (bias) acc += bias[j];
    if (flags & A3D_EPI_RELU) acc = fmaxf(acc, 0.f);
    if (y_f32) reinterpret_cast<float*>(y)[(size_t)m * ldy + j] = acc;
    else reinterpret_cast<uint16_t*>(y)[(size_t)m * ldy + j] = f32_to_bf16_bits(acc);
  }
};

struct ConvDgradProb {
  const uint16_t* dy; const uint16_t* w; uint16_t* dx;
  int N, H, W, C, K, R, S, sh, sw, pt, pl, P, Q, ldy;
  __device__ int dimM() const { return N * H * W; }
  __device__ int dimN() const { return C; }
  __device__ int dimK() const { return R * S * K; }
  __device__ float a(int m, int k) const {
    int co = k % K; int t = k / K; int s = t % S; int r = t / S;
    int iw = m % W; int t2 = m / W; int ih = t2 % H; int n = t2 / H;
    int ph = ih + pt - r, qw = iw + pl - s;
    if (ph < 0 || qw < 0 || ph % sh || qw % sw) return 0.f;
    int p = ph / sh, q = qw / sw;
    if (p >= P || q >= Q) return 0.f;
    return bf16_bits_to_f32(dy[(((size_t)n * P + p) * Q + q) * ldy + co]);
  }
  __device__ float b(int ci, int k) const {
    int co = k % K; int t = k / K;   // t = r*S+s
    return bf16_bits_to_f32(w[((size_t)co * (R * S) + t) * C + ci]);
  }
  __device__ void store(int m, int ci, float acc) const { dx[(size_t)m * C + ci] = f32_to_bf16_bits(acc); }
};

struct ConvWgradProb {
  const uint16_t* x; const uint16_t* dy; float* dw;
  int N, H, W, C, K, R, S, sh, sw, pt, pl, P, Q, ldy;
  int dil, pp;
  __device__ int dimM() const { return K; }
  __device__ int dimN() const { return R * S * C; }
  __device__ int dimK() const { return N * P * Q; }
  __device__ float a(int co, int m) const { return bf16_bits_to_f32(dy[(size_t)m * ldy + co]); }
  __device__ float b(int j, int m) const {
    int c = j % C; int t = j / C; int s = t % S; int r = t / S;
    int q = m % Q; int t2 = m / Q; int p = t2 % P; int n = t2 / P;
    int ih = p * sh - pt + r, iw = q * sw - pl + s * dil;
    if (ih < 0 || ih >= H || iw < 0 || iw >= W) return 0.f;
    return bf16_bits_to_f32(x[(((size_t)n * H + ih) * W + iw) * pp + c]);
  }
  __device__ void store(int co, int j, float acc) const { dw[(size_t)co * (R * S * C) + j] = acc; }
};

struct DenseFwdProb {   // y[m][n] = act(sum_k x[m][k] w[n][k] + bias[n]) * mask
  const uint16_t* x; const uint16_t* w; const float* bias; const uint8_t* mask; void* y;
  int M, N, K, ldx, y_f32; float drop_scale; unsigned flags;
  __device__ int dimM() const { return M; }
  __device__ int dimN() const { return N; }
  __device__ int dimK() const { return K; }
  __device__ float a(int m, int k) const { return bf16_bits_to_f32(x[(size_t)m * ldx + k]); }
  __device__ float b(int n, int k) const { return bf16_bits_to_f32(w[(size_t)n * K + k]); }
  __device__ void store(int m, int n, float acc) const {
    if (bias) acc += bias[n];
    if (flags & A3D_EPI_RELU) acc = fmaxf(acc, 0.f);
    if (flags & A3D_EPI_SIGMOID) acc = 1.f / (1.f + expf(-acc));
    if (mask) acc = mask[(size_t)m * N + n] ? acc * drop_scale : 0.f;
    if (y_f32) reinterpret_cast<float*>(y)[(size_t)m * N + n] = acc;
    else reinterpret_cast<uint16_t*>(y)[(size_t)m * N + n] = f32_to_bf16_bits(acc);
  }
};

struct DenseDgradProb {  // dx[m][k] = sum_n dy[m][n] w[n][k]
  const uint16_t* dy; const uint16_t* w; uint16_t* dx; int M, N, K, lddy;
  __device__ int dimM() const { return M; }
  __device__ int dimN() const { return K; }
  __device__ int dimK() const { return N; }
  __device__ float a(int m, int j) const { return bf16_bits_to_f32(dy[(size_t)m * lddy + j]); }
  __device__ float b(int kin, int j) const { return bf16_bits_to_f32(w[(size_t)j * K + kin]); }
  __device__ void store(int m, int kin, float acc) const { dx[(size_t)m * K + kin] = f32_to_bf16_bits(acc); }
};

struct DenseWgradProb {  // dw[n][k] = sum_m dy[m][n] x[m][k]
  const uint16_t* x; const uint16_t* dy; float* dw; int M, N, K, ldx, lddy;
  __device__ int dimM() const { return N; }
  __device__ int dimN() const { return K; }
  __device__ int dimK() const { return M; }
  __device__ float a(int n, int m) const { return bf16_bits_to_f32(dy[(size_t)m * lddy + n]); }
  __device__ float b(int k, int m) const { return bf16_bits_to_f32(x[(size_t)m * ldx + k]); }
  __device__ void store(int n, int k, float acc) const { dw[(size_t)n * K + k] = acc; }
};

template <class Prob>
__global__ void __launch_bounds__(NT) simt_gemm_kernel(Prob p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int M = p.dimM(), N = p.dimN(), K = p.dimK();
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;   // thread owns rows ty*4.., cols tx*4..
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < (TM * TK) / NT; ++i) {
      int idx = tid + i * NT;
      int kk = idx % TK, mm = idx / TK;
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? p.a(m, k) : 0.f;
      int n = n0 + mm;
      Bs[kk][mm] = (n < N && k < K) ? p.b(n, k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) p.store(m, n, acc[i][j]);
    }
}

template <class Prob>
int launch(a3d_ctx* ctx, const Prob& p, long long M, long long N, cudaStream_t st) {
  dim3 grid(ceil_div(M, TM), ceil_div(N, TN));
  simt_gemm_kernel<Prob><<<grid, NT, 0, st>>>(p);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// column sums of a bf16 [rows, ld] matrix (first C columns, C % 8 == 0) -> f32[C]; out must be zeroed.
// 256 threads = RL row-lanes x C8 column-vectors of 8: 16-byte coalesced loads, per-thread f32
// accumulation over the block's row range, smem reduction across row-lanes, one atomic per column.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const uint16_t* __restrict__ a, size_t rows, int C8, int ld8, float* __restrict__ out,
                   size_t rows_per_block) {
  extern __shared__ float red[];                     // [RL][C8*8]
  const int RL = 256 / C8;
  const int rl = threadIdx.x / C8, cv = threadIdx.x % C8;
  const bool active = rl < RL;
  size_t r0 = (size_t)blockIdx.x * rows_per_block;
  size_t r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const uint4* p = reinterpret_cast<const uint4*>(a);
    for (size_t r = r0 + rl; r < r1; r += RL) {
      uint4 v = __ldg(p + r * ld8 + cv);
      const uint32_t* w = &v.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * k] += __uint_as_float(w[k] << 16);
        acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[(size_t)rl * C8 * 8 + cv * 8 + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < RL; ++j) s += red[(size_t)j * C8 * 8 + c];
    atomicAdd(out + c, s);
  }
}

// few rows, many columns (dense layers: rows = batch): one thread per column, coalesced across the warp,
// no atomics.  out is fully overwritten.
__global__ void colsum_bf16_cols_kernel(const uint16_t* __restrict__ a, int rows, int C, int ld, float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += bf16_bits_to_f32(a[(size_t)r * ld + c]);
  out[c] = s;
}

// the same over a batch stored in rank blocks (dp.dense_gather_adam): row r = row r % gb of block r / gb
__global__ void colsum_bf16_cols_grouped_kernel(const uint16_t* __restrict__ a, int rows, int C, int ld, int gb, size_t gs,
                                                float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += bf16_bits_to_f32(a[(size_t)(r / gb) * gs + (size_t)(r % gb) * ld + c]);
  out[c] = s;
}
// generic fallback (any C, any ld)
__global__ void colsum_bf16_scalar_kernel(const uint16_t* __restrict__ a, size_t rows, int C, int ld,
                                          float* __restrict__ out, int rows_per_block) {
  size_t r0 = (size_t)blockIdx.x * rows_per_block;
  size_t r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (size_t r = r0; r < r1; ++r) s += bf16_bits_to_f32(a[r * ld + c]);
    atomicAdd(out + c, s);
  }
}

}  // namespace

int a3d_simt_conv_fwd(a3d_ctx* ctx, const a3d_conv_desc* d, const uint16_t* x, const uint16_t* w, const float* bias,
                      void* y, int y_dtype, unsigned flags, cudaStream_t st) {
  ConvFwdProb p{x, w, bias, y, d->N, d->H, d->W, d->C, d->K, d->R, d->S, d->stride_h, d->stride_w, d->pad_t, d->pad_l,
                d->P, d->Q, d->ldy, y_dtype == A3D_F32, flags, d->dil_w > 1 ? d->dil_w : 1, d->pix_pitch ? d->pix_pitch : d->C};
  return launch(ctx, p, (long long)d->N * d->P * d->Q, d->K, st);
}

int a3d_simt_conv_dgrad(a3d_ctx* ctx, const a3d_conv_desc* d, const uint16_t* dy, const uint16_t* w, uint16_t* dx,
                        cudaStream_t st) {
  ConvDgradProb p{dy, w, dx, d->N, d->H, d->W, d->C, d->K, d->R, d->S, d->stride_h, d->stride_w, d->pad_t, d->pad_l,
                  d->P, d->Q, d->ldy};
  return launch(ctx, p, (long long)d->N * d->H * d->W, d->C, st);
}

int a3d_simt_conv_wgrad(a3d_ctx* ctx, const a3d_conv_desc* d, const uint16_t* x, const uint16_t* dy, float* dw,
                        cudaStream_t st) {
  ConvWgradProb p{x, dy, dw, d->N, d->H, d->W, d->C, d->K, d->R, d->S, d->stride_h, d->stride_w, d->pad_t, d->pad_l,
                  d->P, d->Q, d->ldy, d->dil_w > 1 ? d->dil_w : 1, d->pix_pitch ? d->pix_pitch : d->C};
  return launch(ctx, p, d->K, (long long)d->R * d->S * d->C, st);
}

int a3d_colsum_bf16_grouped(a3d_ctx* ctx, const uint16_t* a, int rows, int C, int ld, int gb, size_t gs, float* out,
                            cudaStream_t st) {
  colsum_bf16_cols_grouped_kernel<<<ceil_div(C, 128), 128, 0, st>>>(a, rows, C, ld, gb, gs, out);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

int a3d_colsum_bf16(a3d_ctx* ctx, const uint16_t* a, size_t rows, int C, int ld, float* out, cudaStream_t st) {
  if (rows <= 512 && C >= 1024) {
    colsum_bf16_cols_kernel<<<ceil_div(C, 128), 128, 0, st>>>(a, (int)rows, C, ld, out);
    A3D_LAUNCH_OK(ctx);
    return 0;
  }
  A3D_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  const bool vec = C % 8 == 0 && ld % 8 == 0 && C / 8 <= 256 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
  if (vec) {
    const int C8 = C / 8, RL = 256 / C8;
    size_t blocks = (rows + (size_t)RL * 8 - 1) / ((size_t)RL * 8);        // >= 8 rows per row-lane
    size_t cap = (size_t)ctx->sm_count * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    size_t rpb = (rows + blocks - 1) / blocks;
    blocks = (rows + rpb - 1) / rpb;
    colsum_bf16_kernel<<<(int)blocks, 256, (size_t)RL * C * sizeof(float), st>>>(a, rows, C8, ld / 8, out, rpb);
  } else {
    int rpb = 256;
    colsum_bf16_scalar_kernel<<<ceil_div((long long)rows, rpb), 128, 0, st>>>(a, rows, C, ld, out, rpb);
  }
  A3D_LAUNCH_OK(ctx);
  return 0;
}

int a3d_simt_dense_fwd(a3d_ctx* ctx, const uint16_t* x, int ldx, const uint16_t* w, const float* bias,
                       const uint8_t* mask, float drop_rate, void* y, int y_dtype, int M, int N, int K, unsigned flags,
                       cudaStream_t st) {
  DenseFwdProb p{x, w, bias, mask, y, M, N, K, ldx, y_dtype == A3D_F32, 1.f / (1.f - drop_rate), flags};
  return launch(ctx, p, M, N, st);
}

int a3d_simt_dense_dgrad(a3d_ctx* ctx, const uint16_t* dy, int lddy, const uint16_t* w, uint16_t* dx, int M, int N,
                         int K, cudaStream_t st) {
  DenseDgradProb p{dy, w, dx, M, N, K, lddy};
  return launch(ctx, p, M, K, st);
}

// Tiny weight matrices (DCNF dense_1 16x128, dense_2 1x16) over a long batch (768 patches): the 64x64-tile kernel would
// run the whole M loop in one or two CTAs (57 / 93 us).  Here the batch is cut into `gridDim.y` slices, every thread owns
// one dw[n][k] of its slice (x row reads coalesced over k, dy broadcast), slices meet with one atomicAdd per output.
__global__ void small_dense_wgrad_kernel(const uint16_t* __restrict__ x, int ldx, const uint16_t* __restrict__ dy, int lddy,
                                         float* __restrict__ dw, int M, int N, int K) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= N * K) return;
  const int n = o / K, k = o - n * K;
  const int per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * per, m1 = min(M, m0 + per);
  float acc = 0.f;
  for (int m = m0; m < m1; ++m)
    acc += bf16_bits_to_f32(__ldg(dy + (size_t)m * lddy + n)) * bf16_bits_to_f32(__ldg(x + (size_t)m * ldx + k));
  atomicAdd(dw + o, acc);
}

int a3d_simt_dense_wgrad(a3d_ctx* ctx, const uint16_t* x, int ldx, const uint16_t* dy, int lddy, float* dw, int M,
                         int N, int K, cudaStream_t st) {
  if ((long long)N * K <= 4096 && M >= 256) {
    A3D_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), st));
    const int bx = (N * K + 127) / 128;
    int slices = ctx->sm_count * 2 / bx;
    if (slices < 1) slices = 1;
    if (slices > M / 16) slices = M / 16;
    small_dense_wgrad_kernel<<<dim3(bx, slices), 128, 0, st>>>(x, ldx, dy, lddy, dw, M, N, K);
    A3D_LAUNCH_OK(ctx);
    return 0;
  }
  DenseWgradProb p{x, dy, dw, M, N, K, ldx, lddy};
  return launch(ctx, p, N, K, st);
}
