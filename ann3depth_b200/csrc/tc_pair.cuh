// tc_pair.cuh -- CTA-pair ("2-SM") variant of the tcgen05 GEMM engine: two CTAs on the two SMs of a TPC compute ONE
// 256 x BN tile with tcgen05.mma.cta_group::2 (K-major A, tiled or im2col; K-major B).
//
// Why: the one-CTA engine sits on the L2->SM operand stream.  Measured against the per-SM ingest
// cap of ~43 B/clk (12.4 TB/s chip-wide, B300 notes): conv2d_3 forward with 128 x 96 tiles needs 28 KB per 1.57 MFLOP
// k-block = 55 FLOP/B -> 680 TFLOP/s cap, measured 640; conv2d_1 forward with 128 x 256 tiles 87 FLOP/B -> 1080 cap,
// measured 1105 (padded problem).  TMA multicast of the weight tile across a cluster (a round-1 kernel, verified on
// hardware in round 2 and then removed) does not move this: a multicast slice is still DELIVERED to every SM of the
// cluster, so the bytes entering each SM stay the same (and the L2 already merges near-simultaneous reads of a line); it
// measured slower (53.6 vs 47.4 us on conv2d_1 forward).
// With cta_group::2 each SM holds only HALF of the weight tile and the tensor cores of both SMs read both halves:
// per SM and k-block 16 KB of A + BN/2 rows of B, i.e. 128 x 256 per SM at 131 FLOP/B and 128 x 192 at 112 FLOP/B.
//
// Protocol (one tile per CTA pair; warp roles as in tc_gemm.cuh):
//   * cluster of 2 along M; CTA rank r owns rows [256 i + 128 r, +128) and loads, per stage, its A box and rows
//     [r BN/2, (r+1) BN/2) of the B tile into ITS OWN shared memory (same stage offsets in both CTAs);
//   * both producers' loads signal the LEADER's (rank 0) full[s] (.cta_group::2 TMA: complete_tx may land on the peer's
//     mbarrier); the leader's producer is the only arrival and expects the bytes of both CTAs;
//   * the leader's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = BN): rows 0..127 accumulate in the
//     leader's TMEM, rows 128..255 in the peer's, same column range.  tcgen05.alloc.cta_group::2 is a COLLECTIVE of the
//     pair (tools/probe/tmem_pair_probe.cu on B200: one warp of each CTA must issue it -- a leader-only alloc blocks --
//     and both get the same address, one reservation on both SMs);
//   * tcgen05.commit.cta_group::2 multicast to empty[s] of BOTH CTAs releases the stage in both rings; a last commit
//     multicast to tmem_full starts both epilogues;
//   * cluster barrier after the mbarrier init and before TMEM dealloc / exit.
// Epilogues: tc::EPI_TMA_F32 / tc::EPI_TMA_BF16 (no split-K) and, BN = 256, the pool-fused tc::EPI_POOL4_BF16.
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
}  // namespace mc

namespace pair {
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive (+ expect tx bytes) on a barrier that may live in the peer CTA
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes)
               : "memory");
}
// TMA loads whose completion is signalled on `bar_cluster_addr` (this CTA's or the peer's mbarrier); data lands in THIS CTA
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c,
                                                        int w, int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_result)), "r"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t cluster_addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(ptx::smem_u32(bar)), "h"(mask)
               : "memory");
}
}  // namespace pair

// Stage ring of the pair kernel: A box (128 rows) + half a B tile per CTA.  NST stages (2..8).
template <int BN_, int KCB_, int NST_>
struct PairCfg {
  static constexpr int BN = BN_, KCB = KCB_, STAGES = NST_;
  static constexpr int KELEMS = KCB_ / 2;
  static constexpr int HALF_ROWS = BN_ / 2;
  static constexpr int A_BYTES = 128 * KCB_;
  static constexpr int B_BYTES = HALF_ROWS * KCB_;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN_ <= 32 ? 32 : BN_ <= 64 ? 64 : BN_ <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = NST_ * STAGE_BYTES + 1024 + 256;
  static_assert(BN_ % 16 == 0 && BN_ >= 32 && BN_ <= 256, "UMMA N (cta_group::2): multiple of 16, <= 256");
  static_assert(HALF_ROWS % 8 == 0, "a B half must be whole 8-row swizzle atoms");
  static_assert(KCB_ == 128 || KCB_ == 64 || KCB_ == 32, "swizzled K-major rows");
  static_assert(NST_ * STAGE_BYTES >= 4 * 8192, "the TMA epilogue stages 4 x 8 KB in the ring");
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
};

template <class C>
__global__ void __launch_bounds__(192, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB /* box = BN/2 rows */,
                 const __grid_constant__ CUtensorMap tmC, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = mc::cluster_ctarank();
  const int m0 = blockIdx.x * 128;                            // grid.x is even: CTAs 2i, 2i+1 form a pair
  const int n0 = blockIdx.y * C::BN;
  const bool live = m0 < p.M;                                 // the odd CTA of the last pair may lie beyond M
  const int m0_load = live ? m0 : 0;
  const int nkb = p.num_kb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);                        // (the leader's are used) one arrival: the leader's producer
      ptx::mbar_init(&empty_bar[s], 1);                       // one multicast commit of the leader's MMA thread
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) pair::tmem_alloc_2cta<C::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  mc::cluster_sync();                                         // both CTAs' barriers and TMEM exist before any signal
  ptx::tc_fence_after_sync();
  // The accumulator address of a cta_group::2 MMA is ONE address valid in both CTAs' tensor memory: everybody uses the
  // leader's allocation (tcgen05.alloc.cta_group::2 reserves the same columns on both SMs); `tmem_own` is only kept
  // for the matching dealloc of this CTA's own allocation.
  const uint32_t tmem_own = *tmem_slot;
  const uint32_t tmem_base = pair::ld_shared_cluster_u32(pair::map_to_cta(ptx::smem_u32(tmem_slot), 0));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int n_img = 0, h0 = 0, w0 = 0;
      if (p.a_mode == A_IM2COL) {
        n_img = m0_load / p.PQ;
        const int rem = m0_load - n_img * p.PQ;
        const int p0 = rem / p.Q, q0 = rem - p0 * p.Q;
        h0 = p.lower_h + p0 * p.sh;
        w0 = p.lower_w + q0 * p.sw;
      }
      const uint32_t full0_base = pair::map_to_cta(ptx::smem_u32(&full_bar[0]), 0);
      for (int i = 0; i < nkb; ++i) {
        const int stage = i % C::STAGES;
        const uint32_t phase = (uint32_t)(i / C::STAGES) & 1u;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);         // own ring: the pair's MMAs have retired this stage
        uint8_t* sA = smem + stage * C::STAGE_BYTES;
        uint8_t* sB = sA + C::A_BYTES;
        // The leader's producer arrives once, expecting the bytes of BOTH CTAs; the peer only sends data: its loads
        // complete_tx on the leader's barrier (they cannot overtake the phase: the peer passed its own empty[stage], i.e.
        // the pair's MMAs of the previous round have retired, and the new phase still lacks the leader's arrival).
        // No remote arrive / cluster-scope release in the loop: a fence per k-block in this thread costs ~1 us.
        const uint32_t full0 = full0_base + (uint32_t)stage * 8u;                        // the leader's full[stage]
        if (crank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
        if (p.a_mode == A_TILED) {
          pair::tma_load_2d_2cta(sA, &tmA, full0, p.a_k0 + i * C::KELEMS, m0_load);
        } else {
          const int tap = i / p.cblocks, cb = i - tap * p.cblocks;
          const int r = tap / p.S, s = tap - r * p.S;
          pair::tma_load_im2col_4d_2cta(sA, &tmA, full0, cb * C::KELEMS, w0, h0, n_img, (uint16_t)(s * p.dil_w),
                                        (uint16_t)r);
        }
        pair::tma_load_2d_2cta(sB, &tmB, full0, i * C::KELEMS, n0 + (int)crank * C::HALF_ROWS);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(256, C::BN, 0, 0);
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
          pair::umma_bf16_2cta(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (uint32_t)((i | k) != 0));
        pair::umma_commit_2cta(&empty_bar[stage], 3);         // frees the stage in both CTAs' rings
      }
      pair::umma_commit_2cta(tmem_full_bar, 3);               // accumulators complete: both epilogues may start
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
    const int quarter = warp & 3;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after_sync();
    // every MMA of the pair has retired, hence every load into either ring has been consumed: the ring is free as staging
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
  mc::cluster_sync();                                         // the peer's tensor cores may still read this CTA's TMEM / smem
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    pair::tmem_dealloc_2cta<C::TMEM_COLS>(tmem_own);
  }
}

}  // namespace tc
