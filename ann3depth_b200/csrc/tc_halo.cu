// tc_halo.cu -- stride-1 convolution (forward, and dgrad as a forward conv of dY) on tcgen05 with
// SHARED-MEMORY HALO REUSE: the input window of a tile is loaded from L2 once and every filter tap is
// served as a row-shifted view of that window.
//
// Formulation.  The input is a zero-padded NHWC tensor viewed as a 2-D matrix Xp[rows = N*Hp*Wp][C].
// With L = (n*Hp + hp)*Wp + wp the top-left-aligned output position,
//     out[L][co] = sum_{r,s,c} Xp[L + r*Wp + s][c] * W[co][r][s][c]
// so for a tile of 128*TM consecutive L the operand rows of tap (r,s) are the tile's rows shifted by
// r*Wp + s.  The tile's window [L0, L0 + 128*TM + (R-1)*Wp + S-1) x 64 channels is loaded with plain
// 2-D TMA (SWIZZLE_128B, 128-byte rows) and the UMMA A-descriptor simply starts (r*Wp+s)*128 bytes
// further in: the hardware swizzle is a function of the absolute shared-memory address, so a start
// that is not 1024-byte aligned addresses the same bytes TMA wrote (probed on B200, tools/shift_probe.py;
// base_offset stays 0).  Positions with hp >= P or wp >= Q are wrap-around garbage and are not stored.
//
// Per CTA: TM accumulators of 128 x BN f32 in TMEM; the weight tile of one (tap, channel block) is
// streamed through a ring and reused by all TM position tiles.  L2 traffic per MAC drops by ~R*S on
// the activation side and by TM on the weight side compared with one im2col TMA load per tap.
//
//   warp 0: TMA producer (window ring, weight ring)      warp 1: TMEM alloc + MMA issue
//   warps 2-5: epilogue (bias / ReLU / ReLU-mask, bf16 or f32 store into the caller's NHWC tensor)
#include "common.cuh"
#include "ptx.cuh"
#include <cuda.h>

namespace halo {

struct Params {
  int rows_total;            // N*Hp*Wp
  int Wp, HpWp;
  int RS, S, cblocks;        // taps, filter width, 64-channel blocks
  int halo_rows;             // rows of the window actually needed: 128*TM + (R-1)*Wp + S-1
  int a_boxes;               // ceil(halo_rows / 64): 64-row TMA boxes per window
  int P, Q, Nimg;            // valid output extents / images
  int Ncols;                 // valid output channels
  int OH, OW, oph, opw;      // output tensor [Nimg][OH][OW][ldo], position (hp+oph, wp+opw)
  long long ldo;
  void* out;
  int out_f32;
  const float* bias;
  unsigned flags;
  const uint16_t* relu_src;  // optional ReluGrad: out = relu_src[same index] > 0 ? acc : 0
};

template <int BN_, int TM_, int ASTAGES_, int AROWS_>
struct Cfg {
  static constexpr int BN = BN_, TM = TM_, ASTAGES = ASTAGES_;
  static constexpr int AROWS = AROWS_;                     // window capacity in rows (multiple of 64)
  static constexpr int A_BYTES = AROWS_ * 128;
  static constexpr int B_BYTES = BN_ * 128;
  static constexpr int BUDGET = 224 * 1024 - 1024 - 256;
  static constexpr int BSTAGES_RAW = (BUDGET - ASTAGES_ * A_BYTES) / B_BYTES;
  static constexpr int BSTAGES = BSTAGES_RAW > 6 ? 6 : (BSTAGES_RAW < 1 ? 1 : BSTAGES_RAW);
  static constexpr bool VALID = BSTAGES_RAW >= 2 && BN_ * TM_ <= 512;
  static constexpr int TMEM_COLS = (BN_ * TM_ <= 32) ? 32 : (BN_ * TM_ <= 64) ? 64 : (BN_ * TM_ <= 128) ? 128
                                   : (BN_ * TM_ <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = ASTAGES_ * A_BYTES + BSTAGES * B_BYTES + 1024 + 256;
  static_assert(BN_ % 16 == 0 && BN_ >= 16 && BN_ <= 256, "UMMA N");
};

template <class C>
__global__ void __launch_bounds__(192, 1)
conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::ASTAGES * C::A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + C::BSTAGES * C::B_BYTES);
  uint64_t* a_empty = a_full + C::ASTAGES;
  uint64_t* b_full = a_empty + C::ASTAGES;
  uint64_t* b_empty = b_full + C::BSTAGES;
  uint64_t* acc_full = b_empty + C::BSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L0 = blockIdx.x * (128 * C::TM);
  const int n0 = blockIdx.y * C::BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < C::ASTAGES; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C::BSTAGES; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int bi = 0;                                     // running weight-stage counter
      for (int cb = 0; cb < p.cblocks; ++cb) {
        const int as = cb % C::ASTAGES;
        ptx::mbar_wait(&a_empty[as], ((cb / C::ASTAGES) & 1) ^ 1);
        ptx::mbar_expect_tx(&a_full[as], p.a_boxes * 8192);
        for (int j = 0; j < p.a_boxes; ++j)
          ptx::tma_load_2d(sA + as * C::A_BYTES + j * 8192, &tmA, &a_full[as], cb * 64, L0 + j * 64);
        for (int tap = 0; tap < p.RS; ++tap, ++bi) {
          const int bs = bi % C::BSTAGES;
          ptx::mbar_wait(&b_empty[bs], ((bi / C::BSTAGES) & 1) ^ 1);
          ptx::mbar_expect_tx(&b_full[bs], C::B_BYTES);
          ptx::tma_load_2d(sB + bs * C::B_BYTES, &tmB, &b_full[bs], (tap * p.cblocks + cb) * 64, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, C::BN, 0, 0);
      int bi = 0;
      for (int cb = 0; cb < p.cblocks; ++cb) {
        const int as = cb % C::ASTAGES;
        ptx::mbar_wait(&a_full[as], (cb / C::ASTAGES) & 1);
        const uint32_t a_base = ptx::smem_u32(sA + as * C::A_BYTES);
        for (int tap = 0; tap < p.RS; ++tap, ++bi) {
          const int bs = bi % C::BSTAGES;
          ptx::mbar_wait(&b_full[bs], (bi / C::BSTAGES) & 1);
          ptx::tc_fence_after_sync();
          const int r = tap / p.S, s = tap - r * p.S;
          const uint32_t shift = (uint32_t)(r * p.Wp + s) * 128u;
          const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(sB + bs * C::B_BYTES), 16, 1024, ptx::LAYOUT_SW128);
#pragma unroll
          for (int t = 0; t < C::TM; ++t) {
            const uint64_t a_desc = ptx::make_smem_desc(a_base + shift + t * (128 * 128), 16, 1024, ptx::LAYOUT_SW128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(tmem_base + t * C::BN, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                             (uint32_t)((cb | tap | k) != 0));
          }
          ptx::umma_commit(&b_empty[bs]);
        }
        ptx::umma_commit(&a_empty[as]);
      }
      ptx::umma_commit(acc_full);
    }
  } else {
    const int quarter = warp & 3;
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after_sync();
#pragma unroll 1
    for (int t = 0; t < C::TM; ++t) {
      const int L = L0 + t * 128 + quarter * 32 + lane;
      const int n = L / p.HpWp;
      const int rem = L - n * p.HpWp;
      const int hp = rem / p.Wp, wp = rem - hp * p.Wp;
      const bool ok = (L < p.rows_total) && (n < p.Nimg) && (hp < p.P) && (wp < p.Q);
      const long long obase = (((long long)n * p.OH + hp + p.oph) * p.OW + (wp + p.opw)) * p.ldo;
#pragma unroll 1
      for (int c0 = 0; c0 < C::BN; c0 += 16) {
        const int col0 = n0 + c0;
        if (col0 >= p.Ncols) break;
        __syncwarp();
        uint32_t rr[16];
        ptx::tmem_ld_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * C::BN + c0), rr);
        ptx::tmem_ld_wait();
        if (ok) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + j < p.Ncols) v[j] += __ldg(p.bias + col0 + j);
          }
          if (p.flags & A3D_EPI_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          const bool full = col0 + 16 <= p.Ncols;
          if (p.relu_src) {
            const uint16_t* rs = p.relu_src + obase + col0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + j < p.Ncols && !(bf16_bits_to_f32(rs[j]) > 0.f)) v[j] = 0.f;
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + obase + col0;
            if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < p.Ncols) o[j] = v[j];
            }
          } else {
            uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + obase + col0;
            if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
              reinterpret_cast<uint4*>(o)[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                          pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              reinterpret_cast<uint4*>(o)[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                                                          pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j < p.Ncols) o[j] = f32_to_bf16_bits(v[j]);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

}  // namespace halo

// ------------------------------------------------------------------------------------------------ host
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tmap2d_sw128(a3d_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15)) {
    a3d_set_error("halo conv: tensor base / pitch not 16-byte aligned");
    return A3D_EINVAL;
  }
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->fn_encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    a3d_set_error("halo conv: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return A3D_ETMAP;
  }
  return 0;
}

template <class C>
int launch(a3d_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const halo::Params& p, int m_tiles, int n_tiles,
           cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    A3D_CHECK_CUDA(cudaFuncSetAttribute(halo::conv_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  halo::conv_kernel<C><<<dim3(m_tiles, n_tiles), 192, C::SMEM_BYTES, st>>>(tmA, tmB, p);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

// NHWC bf16 [N][H][W][C] (channel stride ld) -> zero-padded [N][Hp][Wp][Cp], interior at (pt, pl)
__global__ void pad_copy_kernel(const uint16_t* __restrict__ src, int N, int H, int W, int C, int ld, uint16_t* __restrict__ dst,
                                int Hp, int Wp, int Cp, int pt, int pl) {
  const int Cp8 = Cp / 8;
  size_t total = (size_t)N * Hp * Wp * Cp8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c8 = (int)(i % Cp8);
    size_t t = i / Cp8;
    int wp = (int)(t % Wp);
    t /= Wp;
    int hp = (int)(t % Hp);
    int n = (int)(t / Hp);
    int h = hp - pt, w = wp - pl;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (h >= 0 && h < H && w >= 0 && w < W && c8 * 8 < C) {
      const uint16_t* s = src + (((size_t)n * H + h) * W + w) * ld + c8 * 8;
      if (c8 * 8 + 8 <= C && ((reinterpret_cast<uintptr_t>(s) & 15) == 0)) v = __ldg(reinterpret_cast<const uint4*>(s));
      else {
        uint16_t* pv = reinterpret_cast<uint16_t*>(&v);
        for (int k = 0; k < 8 && c8 * 8 + k < C; ++k) pv[k] = s[k];
      }
    }
    reinterpret_cast<uint4*>(dst)[i] = v;
  }
}

// w[K][RS][C] -> wd[K][RS][Cp] (zero channel padding), optionally flipped + transposed for dgrad:
// flip: wd[ci][RS-1-t][co (padded to Cp)] = w[co][t][ci]
__global__ void repack_filter_kernel(const uint16_t* __restrict__ w, uint16_t* __restrict__ wd, int K, int RS, int C, int Cp,
                                     int flip) {
  // output dims: rows = flip ? C : K ; inner = Cp where the inner source extent is flip ? K : C
  const int rows = flip ? C : K, inner = flip ? K : C;
  size_t total = (size_t)rows * RS * Cp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Cp);
    size_t t2 = i / Cp;
    int t = (int)(t2 % RS);
    int row = (int)(t2 / RS);
    uint16_t v = 0;
    if (c < inner) v = flip ? w[((size_t)c * RS + (RS - 1 - t)) * C + row] : w[((size_t)row * RS + t) * C + c];
    wd[i] = v;
  }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// geometry of the halo formulation for a stride-1 conv described as a *forward* problem:
// input [N,H,W,C] with top/left zero padding (pt, pl), filter RxS, output extents P x Q.
struct HaloGeom { int Hp, Wp, Cp, need_copy; };

static HaloGeom halo_geom(int H, int W, int C, int ld, int R, int S, int pt, int pl, int P, int Q) {
  HaloGeom g;
  g.Hp = P + R - 1;  if (g.Hp < H + pt) g.Hp = H + pt;
  g.Wp = Q + S - 1;  if (g.Wp < W + pl) g.Wp = W + pl;
  g.Cp = (C + 63) / 64 * 64;
  g.need_copy = !(pt == 0 && pl == 0 && g.Hp == H && g.Wp == W && g.Cp == C && ld == C);
  return g;
}

// scratch for one halo convolution: padded activation copy (if needed) + repacked filter (if needed)
size_t a3d_halo_ws_bytes(int N, int H, int W, int C, int ld, int K, int R, int S, int pt, int pl, int P, int Q, int flip) {
  HaloGeom g = halo_geom(H, W, C, ld, R, S, pt, pl, P, Q);
  size_t b = 0;
  if (g.need_copy) b += align256((size_t)N * g.Hp * g.Wp * g.Cp * 2);
  if (flip) b += align256((size_t)K * R * S * g.Cp * 2);          // rows = K (the conv's C), inner = Cp
  else if (g.Cp != C) b += align256((size_t)K * R * S * g.Cp * 2);
  return b;
}

// Forward-style stride-1 convolution through the halo kernel.
//   x: bf16 [N,H,W,C] (channel stride ldx);  w: bf16 filter, [Kout][R*S][Cw] with Cw == Cp (already packed)
//   out: [N,OH,OW,ldo] at spatial offset (oph,opw); valid outputs P x Q.
int a3d_halo_conv_run(a3d_ctx* ctx, const uint16_t* xp, int N, int Hp, int Wp, int Cp, const uint16_t* wpk, int Kout, int R,
                      int S, int P, int Q, void* out, int out_f32, int OH, int OW, int oph, int opw, long long ldo,
                      const float* bias, unsigned flags, const uint16_t* relu_src, cudaStream_t st) {
  const long long rows_total = (long long)N * Hp * Wp;
  if (rows_total > 0x7fffffffLL) { a3d_set_error("halo conv: too many rows"); return A3D_ENOTSUP; }
  const int cblocks = Cp / 64;
  // tile selection: BN covers Kout when possible; TM = 2 when TMEM and shared memory allow
  int bn = Kout <= 16 ? 16 : Kout <= 64 ? 64 : Kout <= 96 ? 96 : Kout <= 128 ? 128
           : Kout <= 192 ? 192 : (Kout % 256 == 0 ? 256 : (Kout % 192 == 0 ? 192 : 128));
  const int extra = (R - 1) * Wp + (S - 1);
  int tm = 2;
  // keep at least ~one wave of CTAs
  long long ctas2 = ((rows_total + 255) / 256) * ((Kout + bn - 1) / bn);
  if (ctas2 < ctx->sm_count * 3 / 4) tm = 1;
  halo::Params p{};
  p.rows_total = (int)rows_total; p.Wp = Wp; p.HpWp = Hp * Wp; p.RS = R * S; p.S = S; p.cblocks = cblocks;
  p.P = P; p.Q = Q; p.Nimg = N; p.Ncols = Kout; p.OH = OH; p.OW = OW; p.oph = oph; p.opw = opw; p.ldo = ldo;
  p.out = out; p.out_f32 = out_f32; p.bias = bias; p.flags = flags; p.relu_src = relu_src;
  CUtensorMap tmA, tmB;
  int rc = tmap2d_sw128(ctx, &tmA, xp, rows_total, Cp, Cp, 64);
  if (rc) return rc;
  rc = tmap2d_sw128(ctx, &tmB, wpk, Kout, (uint64_t)R * S * Cp, (uint64_t)R * S * Cp, bn);
  if (rc) return rc;
  const int n_tiles = (Kout + bn - 1) / bn;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int halo_rows = 128 * tm + extra;
    const int arows = (halo_rows + 63) / 64 * 64;
    p.halo_rows = halo_rows; p.a_boxes = arows / 64;
    const int m_tiles = (int)((rows_total + 128 * tm - 1) / (128 * tm));
    const int ast = cblocks > 1 ? 2 : 1;
    int rcl = A3D_ENOTSUP;
    bool found = false;
    auto try_cfg = [&](auto cfg) {
      using Cg = decltype(cfg);
      if constexpr (Cg::VALID) {
        if (!found && bn == Cg::BN && tm == Cg::TM && ast == Cg::ASTAGES && arows <= Cg::AROWS) {
          found = true;
          rcl = launch<Cg>(ctx, tmA, tmB, p, m_tiles, n_tiles, st);
        }
      }
    };
#define A3D_HALO_BN(BN)                                                                                          \
  try_cfg(halo::Cfg<BN, 2, 2, 256>{}); try_cfg(halo::Cfg<BN, 2, 2, 448>{}); try_cfg(halo::Cfg<BN, 2, 2, 704>{});   \
  try_cfg(halo::Cfg<BN, 2, 1, 448>{}); try_cfg(halo::Cfg<BN, 2, 1, 704>{}); try_cfg(halo::Cfg<BN, 2, 1, 1024>{}); \
  try_cfg(halo::Cfg<BN, 1, 2, 256>{}); try_cfg(halo::Cfg<BN, 1, 2, 448>{}); try_cfg(halo::Cfg<BN, 1, 2, 704>{});   \
  try_cfg(halo::Cfg<BN, 1, 1, 448>{}); try_cfg(halo::Cfg<BN, 1, 1, 704>{}); try_cfg(halo::Cfg<BN, 1, 1, 1024>{});
    A3D_HALO_BN(16) A3D_HALO_BN(64) A3D_HALO_BN(96) A3D_HALO_BN(128) A3D_HALO_BN(192) A3D_HALO_BN(256)
#undef A3D_HALO_BN
    if (found) return rcl;
    if (tm == 2) { tm = 1; continue; }               // retry with one position tile per CTA
    break;
  }
  a3d_set_error("halo conv: no kernel configuration for BN=%d Wp=%d R=%d S=%d cblocks=%d", bn, Wp, R, S, cblocks);
  return A3D_ENOTSUP;
}

int a3d_halo_pad_copy(a3d_ctx* ctx, const uint16_t* src, int N, int H, int W, int C, int ld, uint16_t* dst, int Hp, int Wp,
                      int Cp, int pt, int pl, cudaStream_t st) {
  size_t total = (size_t)N * Hp * Wp * (Cp / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
  pad_copy_kernel<<<(int)blocks, 256, 0, st>>>(src, N, H, W, C, ld, dst, Hp, Wp, Cp, pt, pl);
  A3D_LAUNCH_OK(ctx);
  return 0;
}

int a3d_halo_repack_filter(a3d_ctx* ctx, const uint16_t* w, uint16_t* wd, int K, int RS, int C, int Cp, int flip,
                           cudaStream_t st) {
  size_t total = (size_t)(flip ? C : K) * RS * Cp;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 8) blocks = (size_t)ctx->sm_count * 8;
  repack_filter_kernel<<<(int)blocks, 256, 0, st>>>(w, wd, K, RS, C, Cp, flip);
  A3D_LAUNCH_OK(ctx);
  return 0;
}
