// CTA-pair (cta_group::2) variant of the tcgen05 mainloop of gemm_sm100.cuh.
//
// A cluster of two CTAs (one TPC) computes a 256 x 256 output tile per step: CTA r holds rows [128 r, 128 r + 128) of
// the A block and of the accumulator, and HALF of the B tile (dst rows [128 r, 128 r + 128)); one tcgen05.mma issued
// by the leader CTA multiplies the pair's operands (M = 256).  Per CTA a stage is 16 KB (A) + 16 KB (B half) instead
// of 48 KB, so the ring is 6 deep in the same shared memory (3072 tensor cycles of latency coverage instead of 2048)
// and every B byte is fetched once per pair instead of once per CTA.
//   * both CTAs' TMA loads signal the LEADER's full barrier (cta_group::2 bulk copy, peer bit cleared);
//   * tcgen05.commit multicasts to the empty / accumulator-full barriers of both CTAs;
//   * both CTAs' epilogue warps arrive (remotely for the peer) on the leader's accumulator-empty barrier.
// Work items are decoded per pair; the epilogue policy is the same as in gemm_sm100.cuh.
#pragma once
#include <stdlib.h>
#include "gemm_sm100.cuh"

namespace vtm {
namespace gemm {

struct Cfg2 {
  static constexpr int BN = 256;
  static constexpr uint32_t B_HALF_BYTES = 128 * BK * 2;       // 16 KB
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_HALF_BYTES;
  static constexpr int STAGES = 6;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr size_t SMEM_BYTES = 1024 + static_cast<size_t>(STAGES) * STAGE_BYTES + 256;
};

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_kernel_2cta(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const Work wk /* m_tiles counts 256-row PAIR blocks */, Epi epi) {
  using C = Cfg2;
  constexpr int STAGES = C::STAGES;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);     // leader: one arrive.expect_tx per phase (+ 64 KB of transactions from both CTAs)
      mbar_init(empty_bar(s), 1);    // one multicast commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * EPI_WARPS);   // leader: epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  cluster_sync_all();      // both CTAs' barriers initialised before any remote signal; both resident before the alloc
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr_addr, C::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer (each CTA: its A rows + its half of B; signals the leader) ===============
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = pair; w < wk.total; w += n_pairs) {
        int m_tile, b, nt0, nt1;
        wk.decode(w, &m_tile, &b, &nt0, &nt1);
        for (int nt = nt0; nt < nt1; ++nt) {
          for (int kc = 0; kc < wk.k_chunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
            const uint32_t full_leader = full_bar(stage) & 0xFEFFFFFFu;   // same offset, CTA 0 of the pair
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            tma_load_3d_2cta(sa, &tmap_a, full_leader, kc * BK, m_tile * 256 + static_cast<int>(rank) * 128, b);
            tma_load_3d_2cta(sa + A_BYTES, &tmap_b, full_leader, kc * BK, nt * BN + static_cast<int>(rank) * 128, b);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_ctr = 0;
      for (int w = pair; w < wk.total; w += n_pairs) {
        int m_tile, b, nt0, nt1;
        wk.decode(w, &m_tile, &b, &nt0, &nt1);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const uint32_t as = tile_ctr & 1u;
          const uint32_t aphase = (tile_ctr >> 1) & 1u;
          mbar_wait(tempty_bar(as), aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kc = 0; kc < wk.k_chunks; ++kc) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128_kmajor(sa);
            const uint64_t bdesc = umma_desc_sw128_kmajor(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_f16_2cta(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
            umma_commit_2cta(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit_2cta(tfull_bar(as));
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    uint32_t tile_ctr = 0;
    Epi e = epi;
    // optional per-warp staging area behind the barriers, as in gemm_sm100.cuh
    if constexpr (Epi::SCRATCH_PER_WARP > 0)
      e.set_scratch(bar_base + 256u + static_cast<uint32_t>(warp - 2) * Epi::SCRATCH_PER_WARP);
    for (int w = pair; w < wk.total; w += n_pairs) {
      int m_tile, b, nt0, nt1;
      wk.decode(w, &m_tile, &b, &nt0, &nt1);
      const int m_tile128 = m_tile * 2 + static_cast<int>(rank);   // this CTA's 128-row block
      e.begin(m_tile128, b, row_in_tile);
      for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
        const uint32_t as = tile_ctr & 1u;
        const uint32_t aphase = (tile_ctr >> 1) & 1u;
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * (BN / 2);
        e.tile(taddr, nt * BN + half * (BN / 2), BN / 2);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(tempty_bar(as));
          else mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));
        }
      }
      e.end(m_tile128, b, row_in_tile);
    }
  }

  tc_fence_before();
  cluster_sync_all();      // nobody leaves (or frees TMEM) while the peer may still signal / be read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
  }
}

// Work plan for 256-row pair blocks: `wk.m_tiles` counts pair blocks, the n range of a block is split until every pair has
// about `target_items` items (the pipeline runs straight across item boundaries).
inline void plan_2cta(Work* wk, int M, int N, int K, int B, int sms, int target_items) {
  wk->plan(M, N, K, B, Cfg2::BN, sms / 2, target_items, 1);
  wk->m_tiles = (M + 255) / 256;
  const long long base = static_cast<long long>(wk->m_tiles) * B;
  int sp = 1;
  while (base * sp < static_cast<long long>(target_items) * (sms / 2) && sp * 2 <= wk->n_tiles) sp *= 2;
  wk->tiles_per_split = (wk->n_tiles + sp - 1) / sp;
  wk->n_splits = (wk->n_tiles + wk->tiles_per_split - 1) / wk->tiles_per_split;
  wk->total = static_cast<int>(base * wk->n_splits);
}

template <class Epi>
inline int launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const Work& wk, const Epi& epi, int sms,
                       cudaStream_t stream) {
  using C = Cfg2;
  const size_t smem = C::SMEM_BYTES + static_cast<size_t>(EPI_WARPS) * Epi::SCRATCH_PER_WARP;
  static_assert(C::SMEM_BYTES + static_cast<size_t>(EPI_WARPS) * Epi::SCRATCH_PER_WARP <= 232448, "shared memory of the pair kernel");
  // shared-memory opt-in: once per instantiation and device (immutable afterwards)
  static bool opted[64];
  int dev = 0;
  int rc = cuda_rc(cudaGetDevice(&dev));
  if (rc) return rc;
  if (dev < 0 || dev >= 64 || !opted[dev]) {
    rc = cuda_rc(cudaFuncSetAttribute(gemm_kernel_2cta<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    if (rc) return rc;
    if (dev >= 0 && dev < 64) opted[dev] = true;
  }
  int pairs = sms / 2;
  if (wk.total < pairs) pairs = wk.total;
  gemm_kernel_2cta<Epi><<<2 * pairs, THREADS, smem, stream>>>(ta, tb, wk, epi);
  return launch_rc();
}

// VTM_GEMM_PAIR=0 / 1 overrides the callers' choice between the single-CTA and the CTA-pair mainloop (A/B runs); -1 = unset.
inline int pair_override() {
  static const int v = [] {
    const char* e = getenv("VTM_GEMM_PAIR");
    return e ? (e[0] == '1' ? 1 : 0) : -1;
  }();
  return v;
}

}  // namespace gemm
}  // namespace vtm
