// tc_persist.cuh -- persistent variant of the tcgen05 GEMM engine (K-major A, tiled or im2col; K-major B).
//
// EXPERIMENTAL (A3D_PERSIST=1 adds it to the tuner's candidates; not yet run on hardware -- written at the end of
// round 1 after the GPU budget was spent).  Why: the MSDN tiles have 9..100 k-blocks, so in
// the one-tile-per-CTA kernel (tc_gemm.cuh) a CTA's life is dominated by its prologue and epilogue; three co-resident
// CTAs hide part of that (1.128 -> 1.072 ms on the step).  Here a CTA stays resident and walks over tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ...:
//   * barrier init, TMEM allocation, tensor-map prefetch are paid once per CTA instead of once per tile;
//   * the accumulator is double-buffered in TMEM (columns [0,BN) and [BN,2BN)): the epilogue warps drain tile j
//     while the producer / MMA issuer are already in the main loop of tile j+1;
//   * the epilogue stages through its OWN 32 KB of shared memory (the stage ring belongs to the next tile).
// Protocol (all mbarriers; u = j >> 1 is the use count of accumulator buffer j & 1):
//   full[s] / empty[s]      producer <-> issuer, one running k-block counter across tiles
//   acc_full[b]             issuer -> epilogue: tcgen05.commit after the tile's last MMA; epilogue waits parity u & 1
//   acc_empty[b]            epilogue -> issuer: 4 arrivals (one per epilogue warp, after its last tcgen05.ld of the
//                           buffer); the issuer waits parity (u & 1) ^ 1 -- immediately true on a fresh barrier
// Supported epilogues: tc::EPI_TMA_F32 and tc::EPI_TMA_BF16 without split-K (bias, ReLU; rows / columns beyond M / N
// are clipped by the output tensor map), BN a multiple of 64 (bf16) or 32 (f32).
#pragma once
#include "tc_gemm.cuh"

namespace tc {

template <class C, int NSTAGE>
struct PersistCfg {
  static_assert(!C::A_MN && !C::B_MN && !C::CHUNKED && C::MT == 1, "persistent kernel: K-major swizzled operands, BM = 128");
  static constexpr int STAGES = NSTAGE;
  static constexpr int RING_BYTES = NSTAGE * C::STAGE_BYTES;
  static constexpr int STAGING_BYTES = 4 * 8192;
  static constexpr int SMEM_BYTES = RING_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int ACC_COLS = 2 * C::BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "two accumulators must fit the 512 TMEM columns");
};

template <class C, int NSTAGE>
__global__ void __launch_bounds__(192, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const Params p, const int tiles_m, const int tiles_n) {
  using P = PersistCfg<C, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + P::RING_BYTES;                    // 1024-byte aligned (STAGE_BYTES is a multiple of 1024)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + P::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* acc_full = empty_bar + NSTAGE;                    // [2]
  uint64_t* acc_empty = acc_full + 2;                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = tiles_m * tiles_n;
  const int nkb = p.num_kb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmC);
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&acc_full[b], 1);
      ptx::mbar_init(&acc_empty[b], 4);                       // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<P::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;                                             // running k-block counter across tiles
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int m0 = (t % tiles_m) * 128;
        const int n0 = (t / tiles_m) * C::BN;
        int n_img = 0, h0 = 0, w0 = 0;
        if (p.a_mode == A_IM2COL) {
          n_img = m0 / p.PQ;
          const int rem = m0 - n_img * p.PQ;
          const int p0 = rem / p.Q, q0 = rem - p0 * p.Q;
          h0 = p.lower_h + p0 * p.sh;
          w0 = p.lower_w + q0 * p.sw;
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = it % NSTAGE;
          const uint32_t phase = (uint32_t)(it / NSTAGE) & 1u;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * C::STAGE_BYTES;
          uint8_t* sB = sA + C::A_BYTES;
          ptx::mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          if (p.a_mode == A_TILED) {
            ptx::tma_load_2d(sA, &tmA, &full_bar[stage], p.a_k0 + kb * C::KELEMS, m0);
          } else {
            const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
            const int r = tap / p.S, s = tap - r * p.S;
            ptx::tma_load_im2col_4d(sA, &tmA, &full_bar[stage], cb * C::KELEMS, w0, h0, n_img, (uint16_t)s, (uint16_t)r);
          }
          ptx::tma_load_2d(sB, &tmB, &full_bar[stage], kb * C::KELEMS, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, C::BN, 0, 0);
      constexpr uint32_t k_layout = C::KCB == 128 ? ptx::LAYOUT_SW128 : C::KCB == 64 ? ptx::LAYOUT_SW64 : ptx::LAYOUT_SW32;
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
        const int buf = j & 1;
        const uint32_t use = (uint32_t)(j >> 1);
        ptx::mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);     // the epilogue has drained this buffer's previous tile
        ptx::tc_fence_after_sync();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * C::BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = it % NSTAGE;
          const uint32_t phase = (uint32_t)(it / NSTAGE) & 1u;
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sA = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sB = sA + C::A_BYTES;
          const uint64_t a_desc = ptx::make_smem_desc(sA, 16, 8 * C::KCB, k_layout);
          const uint64_t b_desc = ptx::make_smem_desc(sB, 16, 8 * C::KCB, k_layout);
#pragma unroll
          for (int k = 0; k < C::KELEMS / 16; ++k)
            ptx::umma_bf16(tmem_acc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                           (uint32_t)((kb | k) != 0));
          ptx::umma_commit(&empty_bar[stage]);                // frees the stage when these MMAs retire
        }
        ptx::umma_commit(&acc_full[buf]);                     // this tile's accumulator is complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
    uint8_t* slab0 = staging + quarter * 8192;                // two 4 KB slabs per warp
    int j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int buf = j & 1;
      const uint32_t use = (uint32_t)(j >> 1);
      const int m0 = (t % tiles_m) * 128;
      const int n0 = (t / tiles_m) * C::BN;
      ptx::mbar_wait(&acc_full[buf], use & 1u);
      ptx::tc_fence_after_sync();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * C::BN) + ((uint32_t)(quarter * 32) << 16);
      const int row0 = m0 + quarter * 32;
      const bool rows_live = row0 < p.M;                      // warp-uniform
      if (p.epi == EPI_TMA_F32) {
        constexpr int NCH = C::BN / 32;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const int col0 = n0 + ch * 32;
          uint32_t ra[16], rb[16];
          __syncwarp();
          ptx::tmem_ld_x16(tmem_acc + (uint32_t)(ch * 32), ra);
          ptx::tmem_ld_x16(tmem_acc + (uint32_t)(ch * 32 + 16), rb);
          ptx::tmem_ld_wait();
          if (!rows_live || col0 >= p.N) continue;            // still read: keeps the loop warp-uniform and simple
          uint8_t* slab = slab0 + (ch & 1) * 4096;
          if (ch >= 2) {
            if (lane == 0) ptx::bulk_wait_read<1>();          // the store issued two chunks ago has read this slab
          }
          __syncwarp();
          float v[32];
#pragma unroll
          for (int q = 0; q < 16; ++q) { v[q] = __uint_as_float(ra[q]); v[16 + q] = __uint_as_float(rb[q]); }
          if (p.bias) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (col0 + q < p.N) v[q] += __ldg(p.bias + col0 + q);
          }
          if (p.flags & A3D_EPI_RELU) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
          }
          const uint32_t srow = ptx::smem_u32(slab) + (uint32_t)lane * 128u;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            ptx::st_shared_v4(srow + (uint32_t)((q ^ (lane & 7)) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmC, slab, col0, row0);
            ptx::bulk_commit();
          }
        }
      } else {                                                // EPI_TMA_BF16: 64 bf16 columns (128 bytes) per slab row
        constexpr int NCH = C::BN / 64;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const int col0 = n0 + ch * 64;
          uint8_t* slab = slab0 + (ch & 1) * 4096;
          const bool live = rows_live && col0 < p.N;          // warp-uniform
          if (live && ch >= 2) {
            if (lane == 0) ptx::bulk_wait_read<1>();
          }
          __syncwarp();
          const uint32_t srow = ptx::smem_u32(slab) + (uint32_t)lane * 128u;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t ra[16], rb[16];
            ptx::tmem_ld_x16(tmem_acc + (uint32_t)(ch * 64 + half * 32), ra);
            ptx::tmem_ld_x16(tmem_acc + (uint32_t)(ch * 64 + half * 32 + 16), rb);
            ptx::tmem_ld_wait();
            if (!live) continue;
            float v[32];
#pragma unroll
            for (int q = 0; q < 16; ++q) { v[q] = __uint_as_float(ra[q]); v[16 + q] = __uint_as_float(rb[q]); }
            const int cbase = col0 + half * 32;
            if (p.bias) {
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (cbase + q < p.N) v[q] += __ldg(p.bias + cbase + q);
            }
            if (p.flags & A3D_EPI_RELU) {
#pragma unroll
              for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
            }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const int q = half * 4 + qq;
              ptx::st_shared_v4_b32(srow + (uint32_t)((q ^ (lane & 7)) << 4), pack_bf16x2(v[8 * qq], v[8 * qq + 1]),
                                    pack_bf16x2(v[8 * qq + 2], v[8 * qq + 3]), pack_bf16x2(v[8 * qq + 4], v[8 * qq + 5]),
                                    pack_bf16x2(v[8 * qq + 6], v[8 * qq + 7]));
            }
          }
          if (!live) continue;
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmC, slab, col0, row0);
            ptx::bulk_commit();
          }
        }
      }
      // every tcgen05.ld of this buffer has completed (wait::ld above): hand it back to the issuer
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&acc_empty[buf]);
        ptx::bulk_wait_read<0>();                             // the slabs are free for the next tile
      }
      __syncwarp();
    }
    if (lane == 0) ptx::bulk_wait_all();                      // all stores of this warp are complete before exit
    __syncwarp();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<P::TMEM_COLS>(tmem_base);
  }
}

}  // namespace tc
