// tc_mcast.cuh -- cluster variant of the tcgen05 GEMM engine: the B (weight) tile is loaded ONCE per cluster and
// multicast to all its CTAs (K-major A, tiled or im2col; K-major B).
//
// EXPERIMENTAL (A3D_MCAST=1 adds it to the tuner's candidates; written at the end of round 1 after the GPU budget was
// spent -- compiled, not yet run on hardware).  Why: ncu on conv2d_1 forward (128x256 tiles)
// shows the MMA issuer starved (it spins on the full barriers) while the producer never waits for a free stage: the
// operand stream itself is the limit.  A 128x256x64 k-block needs 48 KB for 4.2 MFLOP = 87 FLOP/B; the chip-wide
// L2->SM cap of ~6300 B/clk is 42.6 B/clk per SM, i.e. 3.7 kFLOP/clk/SM = 45 % of the tensor pipe -- the measured
// 49 %.  Two thirds of those bytes are the weight tile, which is the same for every M tile.  With a cluster of CL CTAs
// along M, CTA r loads rows [r*BN/CL, (r+1)*BN/CL) of each B stage and multicasts them: 16 + 32/CL KB per CTA and
// k-block (CL = 2: 131 FLOP/B, CL = 4: 175 FLOP/B).
//
// Protocol differences to tc_gemm.cuh (one tile per CTA, same warp roles):
//   * full[s] of every CTA still expects STAGE_BYTES: its own A box + CL multicast slices of B;
//   * a slice may only be multicast into stage s once EVERY CTA of the cluster has retired the MMAs that read s:
//     empty[s] counts CL arrivals, and each issuer's tcgen05.commit is multicast to the empty[s] of all CTAs;
//   * cluster barrier after the mbarrier init (no remote arrive / multicast may reach an uninitialised barrier) and
//     before exit (no CTA leaves while peers can still signal its barriers);
//   * grid.x is padded to a multiple of CL; a padding CTA (m0 >= M) runs the whole protocol on tile 0's A operand and
//     stores nothing.
// Epilogues: tc::EPI_TMA_F32 / tc::EPI_TMA_BF16 without split-K, as in tc_persist.cuh, and (BN = 256) the pool-fused
// tc::EPI_POOL4_BF16 of fine/first.
#pragma once
#include "tc_gemm.cuh"

namespace tc {

namespace mc {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D tiled load multicast to the CTAs in `mask`: the box lands at the same CTA-relative shared-memory offset in each
// of them and completes tx bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1),
      "h"(mask)
      : "memory");
}
// arrive (once all previously issued UMMAs of this thread have completed) on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(ptx::smem_u32(bar)), "h"(mask)
               : "memory");
}
}  // namespace mc

template <class C, int CL>
__global__ void __launch_bounds__(192, 1)
gemm_mcast_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB /* box = BN/CL rows */,
                  const __grid_constant__ CUtensorMap tmC, const Params p) {
  static_assert(!C::A_MN && !C::B_MN && !C::CHUNKED && C::MT == 1, "multicast kernel: K-major swizzled operands, BM = 128");
  static_assert(CL == 2 || CL == 4, "cluster of 2 or 4 CTAs along M");
  static_assert((C::BN / CL) % 8 == 0 && C::BN % CL == 0, "a B slice must be whole 8-row swizzle atoms");
  constexpr int SLICE_ROWS = C::BN / CL;
  constexpr int SLICE_BYTES = SLICE_ROWS * C::KCB;
  constexpr uint16_t MASK = (uint16_t)((1u << CL) - 1u);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = mc::cluster_ctarank();
  const int m0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * C::BN;
  const bool live = m0 < p.M;                                 // padding CTAs of the last cluster store nothing
  const int m0_load = live ? m0 : 0;
  const int nkb = p.num_kb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], CL);                      // one arrival per CTA of the cluster
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  mc::cluster_sync();                                         // every CTA's barriers exist before anybody signals them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int n_img = 0, h0 = 0, w0 = 0;
      if (p.a_mode == A_IM2COL) {
        n_img = m0_load / p.PQ;
        const int rem = m0_load - n_img * p.PQ;
        const int p0 = rem / p.Q, q0 = rem - p0 * p.Q;
        h0 = p.lower_h + p0 * p.sh;
        w0 = p.lower_w + q0 * p.sw;
      }
      for (int i = 0; i < nkb; ++i) {
        const int stage = i % C::STAGES;
        const uint32_t phase = (uint32_t)(i / C::STAGES) & 1u;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);         // all CL CTAs have released this stage
        uint8_t* sA = smem + stage * C::STAGE_BYTES;
        uint8_t* sB = sA + C::A_BYTES;
        ptx::mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
        if (p.a_mode == A_TILED) {
          ptx::tma_load_2d(sA, &tmA, &full_bar[stage], p.a_k0 + i * C::KELEMS, m0_load);
        } else {
          const int tap = i / p.cblocks, cb = i - tap * p.cblocks;
          const int r = tap / p.S, s = tap - r * p.S;
          ptx::tma_load_im2col_4d(sA, &tmA, &full_bar[stage], cb * C::KELEMS, w0, h0, n_img, (uint16_t)s, (uint16_t)r);
        }
        // this CTA's slice of the weight tile, to every CTA of the cluster
        mc::tma_load_2d_mcast(sB + crank * SLICE_BYTES, &tmB, &full_bar[stage], i * C::KELEMS,
                              n0 + (int)crank * SLICE_ROWS, MASK);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, C::BN, 0, 0);
      constexpr uint32_t k_layout = C::KCB == 128 ? ptx::LAYOUT_SW128 : C::KCB == 64 ? ptx::LAYOUT_SW64 : ptx::LAYOUT_SW32;
      for (int i = 0; i < nkb; ++i) {
        const int stage = i % C::STAGES;
        const uint32_t phase = (uint32_t)(i / C::STAGES) & 1u;
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after_sync();
        const uint32_t sA = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
        const uint32_t sB = sA + C::A_BYTES;
        const uint64_t a_desc = ptx::make_smem_desc(sA, 16, 8 * C::KCB, k_layout);
        const uint64_t b_desc = ptx::make_smem_desc(sB, 16, 8 * C::KCB, k_layout);
#pragma unroll
        for (int k = 0; k < C::KELEMS / 16; ++k)
          ptx::umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (uint32_t)((i | k) != 0));
        mc::umma_commit_mcast(&empty_bar[stage], MASK);       // this CTA is done with the stage: tell every producer
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after_sync();
    // All loads of THIS CTA have landed and all its MMAs have read their stages, but peers may still multicast later
    // k-blocks?  No: every CTA runs the same nkb rounds and a round's slices are only sent once all CTAs released the
    // stage, so after this CTA's last full barrier completed nothing more is written into its ring: it is free as staging.
    uint8_t* slab0 = smem + quarter * 8192;
    const int row0 = m0 + quarter * 32;
    const uint32_t tmem_acc = tmem_base + ((uint32_t)(quarter * 32) << 16);
    if (live && row0 < p.M) {                                 // warp-uniform
      if (p.epi == EPI_TMA_F32) {
        constexpr int NCH = C::BN / 32;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const int col0 = n0 + ch * 32;
          if (col0 >= p.N) break;
          uint8_t* slab = slab0 + (ch & 1) * 4096;
          if (ch >= 2) {
            if (lane == 0) ptx::bulk_wait_read<1>();
          }
          __syncwarp();
          uint32_t ra[16], rb[16];
          ptx::tmem_ld_x16(tmem_acc + (uint32_t)(ch * 32), ra);
          ptx::tmem_ld_x16(tmem_acc + (uint32_t)(ch * 32 + 16), rb);
          ptx::tmem_ld_wait();
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
      } else if (C::BN == 256 && p.epi == EPI_POOL4_BF16) {
        // fused 2x2 max-pool over the four 64-column groups (same epilogue as tc_gemm.cuh EPI_POOL4_BF16)
        uint8_t* slab = slab0;
        const int row = row0 + lane;
        const bool row_ok = row < p.M;
        const uint32_t srow = ptx::smem_u32(slab) + (uint32_t)lane * 128u;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t r0[16], r1[16], r2[16], r3[16];
          const uint32_t taddr = tmem_acc + (uint32_t)(cc * 16);
          __syncwarp();
          ptx::tmem_ld_x16(taddr, r0);
          ptx::tmem_ld_x16(taddr + 64, r1);
          ptx::tmem_ld_x16(taddr + 128, r2);
          ptx::tmem_ld_x16(taddr + 192, r3);
          ptx::tmem_ld_wait();
          float v[16];
          uint32_t gi[4] = {0, 0, 0, 0};
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float m = __uint_as_float(r0[q]);
            uint32_t g = 0;
            const float a1 = __uint_as_float(r1[q]), a2 = __uint_as_float(r2[q]), a3 = __uint_as_float(r3[q]);
            if (a1 > m) { m = a1; g = 1; }                    // strict '>' keeps the FIRST arg-max (TF MaxPoolGrad)
            if (a2 > m) { m = a2; g = 2; }
            if (a3 > m) { m = a3; g = 3; }
            if (p.bias) m += __ldg(p.bias + cc * 16 + q);
            if (p.flags & A3D_EPI_RELU) m = fmaxf(m, 0.f);
            v[q] = m;
            gi[q >> 2] |= g << ((q & 3) * 8);
          }
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) {
            const int q = cc * 2 + qq;
            ptx::st_shared_v4_b32(srow + (uint32_t)((q ^ (lane & 7)) << 4), pack_bf16x2(v[8 * qq], v[8 * qq + 1]),
                                  pack_bf16x2(v[8 * qq + 2], v[8 * qq + 3]), pack_bf16x2(v[8 * qq + 4], v[8 * qq + 5]),
                                  pack_bf16x2(v[8 * qq + 6], v[8 * qq + 7]));
          }
          if (p.pool_idx && row_ok)
            *reinterpret_cast<uint4*>(p.pool_idx + (size_t)row * 64 + cc * 16) = make_uint4(gi[0], gi[1], gi[2], gi[3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tmC, slab, 0, row0);
          ptx::bulk_commit();
        }
      } else {                                                // EPI_TMA_BF16
        constexpr int NCH = C::BN / 64;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const int col0 = n0 + ch * 64;
          if (col0 >= p.N) break;
          uint8_t* slab = slab0 + (ch & 1) * 4096;
          if (ch >= 2) {
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
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmC, slab, col0, row0);
            ptx::bulk_commit();
          }
        }
      }
      if (lane == 0) ptx::bulk_wait_all();
      __syncwarp();
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
  mc::cluster_sync();                                         // peers may still arrive on this CTA's empty barriers
}

}  // namespace tc
