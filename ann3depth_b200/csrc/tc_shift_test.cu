// tc_shift_test.cu -- hardware experiment (unit-test hook, not on the product path): can a UMMA
// shared-memory descriptor start at an arbitrary 128-byte ROW offset inside a SWIZZLE_128B tile?
// This decides whether convolution taps can be served as row-shifted views of one halo tile that
// TMA loaded once (instead of one TMA load per tap).  The descriptor's `base_offset` field
// (bits 49-51) is documented as (start_address >> 7) & 7 for starts that are not 1024-byte aligned.
#include "common.cuh"
#include "ptx.cuh"
#include <cuda.h>

namespace {

// D[128][N] = A[shift : shift+128][0:64] * B[N][64]^T      (a_mn = 0, A K-major  [rows][64])
// D[128][N] = A[shift : shift+64 ][0:128]^T * B[N][64]^T   (a_mn = 1, A MN-major [K rows][128])
__global__ void __launch_bounds__(192, 1)
shift_test_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int N,
                  int shift, int use_base_offset, int a_mn) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                        // K-major: 136 rows x 128 B ; MN-major: 2 blocks of 72 rows x 128 B
  uint8_t* sB = smem + 20 * 1024;            // N rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20 * 1024 + 32 * 1024);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {
    if (!a_mn) {
      ptx::mbar_expect_tx(bar, 136 * 128 + N * 128);
      ptx::tma_load_2d(sA, &tmA, bar, 0, 0);                       // box 64 cols x 136 rows
    } else {
      ptx::mbar_expect_tx(bar, 2 * 72 * 128 + N * 128);
      ptx::tma_load_2d(sA, &tmA, bar, 0, 0);                       // box 64 cols x 72 rows (cols 0..63)
      ptx::tma_load_2d(sA + 72 * 128, &tmA, bar, 64, 0);           // cols 64..127
    }
    ptx::tma_load_2d(sB, &tmB, bar, 0, 0);
  } else if (warp == 1 && lane == 0) {
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after_sync();
    const uint32_t a_addr = ptx::smem_u32(sA) + shift * 128;
    uint64_t a_desc = a_mn ? ptx::make_smem_desc(a_addr, 72 * 128, 1024, ptx::LAYOUT_SW128)
                           : ptx::make_smem_desc(a_addr, 16, 1024, ptx::LAYOUT_SW128);
    if (use_base_offset) a_desc |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t b_desc = ptx::make_smem_desc(ptx::smem_u32(sB), 16, 1024, ptx::LAYOUT_SW128);
    const uint32_t idesc = ptx::make_idesc_bf16(128, N, a_mn, 0);
    const uint32_t a_step = a_mn ? (2048 >> 4) : 2;
    for (int k = 0; k < 4; ++k)
      ptx::umma_bf16(tmem_base, a_desc + (uint64_t)(k * a_step), b_desc + (uint64_t)(k * 2), idesc, k != 0);
    ptx::umma_commit(done);
  } else if (warp >= 2) {
    const int quarter = warp & 3;
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after_sync();
    const int row = quarter * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      __syncwarp();
      uint32_t r[16];
      ptx::tmem_ld_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + c0, r);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<256>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tmap2d(a3d_ctx* ctx, CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->fn_encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { a3d_set_error("shift test: tensor map encode failed (%d)", (int)r); return A3D_ETMAP; }
  return 0;
}

}  // namespace

// A: a_mn=0 -> bf16 [136][64]; a_mn=1 -> bf16 [72][128].  B: bf16 [N][64].  D: f32 [128][N].
extern "C" int a3d_debug_tc_shift(a3d_ctx* ctx, const uint16_t* A, const uint16_t* B, float* D, int N, int shift,
                                  int use_base_offset, int a_mn, void* stream) {
  A3D_REQUIRE(ctx && A && B && D && N % 16 == 0 && N >= 16 && N <= 256 && shift >= 0 && shift <= 8, "shift test: bad arg");
  CUtensorMap tmA, tmB;
  int rc = a_mn ? tmap2d(ctx, &tmA, A, 72, 128, 72) : tmap2d(ctx, &tmA, A, 136, 64, 136);
  if (rc) return rc;
  rc = tmap2d(ctx, &tmB, B, N, 64, N);
  if (rc) return rc;
  const int smem = 20 * 1024 + 32 * 1024 + 1024 + 256;
  A3D_CHECK_CUDA(cudaFuncSetAttribute(shift_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  shift_test_kernel<<<1, 192, smem, as_stream(stream)>>>(tmA, tmB, D, N, shift, use_base_offset, a_mn);
  A3D_LAUNCH_OK(ctx);
  return 0;
}
