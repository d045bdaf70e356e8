// conv.cu -- public convolution / dense entry points: argument checks and implementation dispatch
// (tcgen05 engine in tc_gemm.cu where the shape allows, CUDA-core implicit GEMM otherwise).
#include "common.cuh"

int a3d_tc_conv_fwd(a3d_ctx*, const a3d_conv_desc*, const uint16_t* x, const uint16_t* w, const float* bias, void* y,
                    int y_dtype, unsigned flags, void* ws, size_t ws_bytes, cudaStream_t st);
size_t a3d_tc_conv_fwd_ws_bytes(a3d_ctx*, const a3d_conv_desc*);
int a3d_tc_dense_fwd(a3d_ctx*, const uint16_t* x, int ldx, const uint16_t* w, const float* bias, const uint8_t* mask,
                     float drop_rate, void* y, int y_dtype, float* acc_ws, int M, int N, int K, unsigned flags,
                     cudaStream_t st);

static int check_desc(const a3d_conv_desc* d) {
  A3D_REQUIRE(d, "conv: null descriptor");
  A3D_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->K > 0 && d->R > 0 && d->S > 0, "conv: bad dims");
  A3D_REQUIRE(d->stride_h > 0 && d->stride_w > 0 && d->pad_t >= 0 && d->pad_l >= 0, "conv: bad stride/pad");
  A3D_REQUIRE(d->P > 0 && d->Q > 0 && d->ldy >= d->K, "conv: bad output dims");
  // every tap of every output pixel must start inside the padded input
  A3D_REQUIRE((d->P - 1) * d->stride_h - d->pad_t < d->H && (d->Q - 1) * d->stride_w - d->pad_l < d->W,
              "conv: output larger than the input supports");
  return 0;
}

// dgrad of a stride-1 conv == forward conv of dy with the spatially flipped, channel-transposed filter
static bool dgrad_as_fwd_ok(const a3d_conv_desc* d) {
  return d->stride_h == 1 && d->stride_w == 1 && d->ldy == d->K && d->K % 16 == 0 && d->C % 8 == 0 &&
         d->R - 1 - d->pad_t >= 0 && d->S - 1 - d->pad_l >= 0;
}

__global__ void flip_filter_kernel(const uint16_t* __restrict__ w, uint16_t* __restrict__ wd, int K, int RS, int C) {
  // w[co][t][ci] -> wd[ci][RS-1-t][co]
  size_t total = (size_t)K * RS * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int co = (int)(i % K);
    size_t t2 = i / K;
    int tf = (int)(t2 % RS);
    int ci = (int)(t2 / RS);
    wd[i] = w[((size_t)co * RS + (RS - 1 - tf)) * C + ci];
  }
}

extern "C" size_t a3d_conv2d_ws_bytes(a3d_ctx* ctx, const a3d_conv_desc* d, int op) {
  if (!d) return 0;
  size_t filt = ((size_t)d->K * d->R * d->S * d->C * 2 + 255) & ~(size_t)255;
  if (op == A3D_OP_FWD) return (size_t)d->N * d->P * d->Q * d->K * sizeof(float);
  if (op == A3D_OP_DGRAD) return filt + (size_t)d->N * d->H * d->W * d->C * sizeof(float);
  return 0;
}

extern "C" int a3d_conv2d_fwd(a3d_ctx* ctx, const a3d_conv_desc* d, const uint16_t* x, const uint16_t* w,
                              const float* bias, void* y, int y_dtype, unsigned flags, void* ws, size_t ws_bytes,
                              void* stream) {
  A3D_REQUIRE(ctx && x && w && y, "conv fwd: null argument");
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (d->impl == A3D_IMPL_SIMT) return a3d_simt_conv_fwd(ctx, d, x, w, bias, y, y_dtype, flags, st);
  if (a3d_tc_conv_fwd_supported(d)) return a3d_tc_conv_fwd(ctx, d, x, w, bias, y, y_dtype, flags, ws, ws_bytes, st);
  if (d->impl == A3D_IMPL_TC) {
    a3d_set_error("conv fwd: shape not supported by the tcgen05 path (C=%d K=%d)", d->C, d->K);
    return A3D_ENOTSUP;
  }
  return a3d_simt_conv_fwd(ctx, d, x, w, bias, y, y_dtype, flags, st);
}

extern "C" int a3d_conv2d_dgrad(a3d_ctx* ctx, const a3d_conv_desc* d, const uint16_t* dy, const uint16_t* w,
                                uint16_t* dx, void* ws, size_t ws_bytes, void* stream) {
  A3D_REQUIRE(ctx && dy && w && dx, "conv dgrad: null argument");
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  size_t filt = ((size_t)d->K * d->R * d->S * d->C * 2 + 255) & ~(size_t)255;
  if (d->impl != A3D_IMPL_SIMT && dgrad_as_fwd_ok(d) && ws && ws_bytes >= filt) {
    uint16_t* wd = reinterpret_cast<uint16_t*>(ws);
    size_t total = (size_t)d->K * d->R * d->S * d->C;
    int grid = (int)((total + 255) / 256);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    flip_filter_kernel<<<grid, 256, 0, st>>>(w, wd, d->K, d->R * d->S, d->C);
    A3D_LAUNCH_OK(ctx);
    a3d_conv_desc e = *d;
    e.H = d->P; e.W = d->Q; e.C = d->K;          // "input" is dy
    e.K = d->C; e.P = d->H; e.Q = d->W; e.ldy = d->C;
    e.pad_t = d->R - 1 - d->pad_t; e.pad_l = d->S - 1 - d->pad_l;
    e.stride_h = e.stride_w = 1;
    return a3d_tc_conv_fwd(ctx, &e, dy, wd, nullptr, dx, A3D_BF16, 0, reinterpret_cast<uint8_t*>(ws) + filt,
                           ws_bytes - filt, st);
  }
  if (d->impl == A3D_IMPL_TC) {
    a3d_set_error("conv dgrad: shape not supported by the tcgen05 path (stride %d, K=%d, ws=%zu)", d->stride_h, d->K,
                  ws_bytes);
    return A3D_ENOTSUP;
  }
  return a3d_simt_conv_dgrad(ctx, d, dy, w, dx, st);
}

extern "C" int a3d_conv2d_wgrad(a3d_ctx* ctx, const a3d_conv_desc* d, const uint16_t* x, const uint16_t* dy, float* dw,
                                float* db, void* ws, size_t ws_bytes, void* stream) {
  A3D_REQUIRE(ctx && x && dy && dw, "conv wgrad: null argument");
  int rc = check_desc(d);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (db) {
    rc = a3d_colsum_bf16(ctx, dy, (size_t)d->N * d->P * d->Q, d->K, d->ldy, db, st);
    if (rc) return rc;
  }
  if (d->impl == A3D_IMPL_TC) {
    a3d_set_error("conv wgrad: tcgen05 path not available for this shape");
    return A3D_ENOTSUP;
  }
  return a3d_simt_conv_wgrad(ctx, d, x, dy, dw, st);
}

// ------------------------------------------------------------------------------------------ dense
extern "C" int a3d_dense_fwd(a3d_ctx* ctx, const uint16_t* x, int ldx, const uint16_t* w, const float* bias,
                             const uint8_t* keep_mask, float drop_rate, void* y, int y_dtype, float* acc_ws, int M,
                             int N, int K, unsigned flags, int impl, void* stream) {
  A3D_REQUIRE(ctx && x && w && y && M > 0 && N > 0 && K > 0 && ldx >= K, "dense fwd: bad argument");
  cudaStream_t st = as_stream(stream);
  bool tc_ok = (K % 64 == 0) && (ldx % 8 == 0) && M <= 256 && acc_ws;
  if (impl != A3D_IMPL_SIMT && tc_ok)
    return a3d_tc_dense_fwd(ctx, x, ldx, w, bias, keep_mask, drop_rate, y, y_dtype, acc_ws, M, N, K, flags, st);
  if (impl == A3D_IMPL_TC) {
    a3d_set_error("dense fwd: shape not supported by the tcgen05 path (M=%d N=%d K=%d)", M, N, K);
    return A3D_ENOTSUP;
  }
  return a3d_simt_dense_fwd(ctx, x, ldx, w, bias, keep_mask, drop_rate, y, y_dtype, M, N, K, flags, st);
}

extern "C" int a3d_dense_dgrad(a3d_ctx* ctx, const uint16_t* dy, const uint16_t* w, uint16_t* dx, float* acc_ws, int M,
                               int N, int K, int impl, void* stream) {
  A3D_REQUIRE(ctx && dy && w && dx && M > 0 && N > 0 && K > 0, "dense dgrad: bad argument");
  if (impl == A3D_IMPL_TC) {
    a3d_set_error("dense dgrad: tcgen05 path not available");
    return A3D_ENOTSUP;
  }
  return a3d_simt_dense_dgrad(ctx, dy, w, dx, M, N, K, as_stream(stream));
}

extern "C" int a3d_dense_wgrad(a3d_ctx* ctx, const uint16_t* x, int ldx, const uint16_t* dy, float* dw, float* db, int M,
                               int N, int K, int impl, void* stream) {
  A3D_REQUIRE(ctx && x && dy && dw && M > 0 && N > 0 && K > 0, "dense wgrad: bad argument");
  cudaStream_t st = as_stream(stream);
  if (db) {
    int rc = a3d_colsum_bf16(ctx, dy, (size_t)M, N, N, db, st);
    if (rc) return rc;
  }
  if (impl == A3D_IMPL_TC) {
    a3d_set_error("dense wgrad: tcgen05 path not available");
    return A3D_ENOTSUP;
  }
  return a3d_simt_dense_wgrad(ctx, x, ldx, dy, dw, M, N, K, st);
}
