// Persistent warp-specialised tcgen05 GEMM mainloop for sm_100a, shared by KA (similarity + row
// arg-max epilogue) and the projection GEMMs of KD (fp16 store epilogue).
//
//   D[b][m, n] = sum_k A[b][m, k] * Bm[b][n, k]        fp16 operands (K-major), fp32 accumulators
//
// One CTA = 320 threads: warp 0 TMA producer (one lane), warp 1 TMEM allocator + MMA issuer (one lane),
// warps 2..9 epilogue (TMEM lane quadrant = warp_id % 4; column half = (warp_id - 2) / 4).
// Shared memory: STAGES x (A tile 128x64 + B tile BNx64) fp16, 128-byte swizzled by TMA, consumed in
// place by tcgen05.mma through shared-memory descriptors.  TMEM: two accumulators of BN fp32 columns
// (double buffered) so the epilogue of tile i overlaps the MMAs of tile i+1.
// A work item is (m_tile, b, [nt0, nt1)); CTAs stride over work items (persistent grid = #SMs).
//
// Epilogue policy `Epi` (all members __device__, called by every epilogue thread):
//   void begin(int m_tile, int b, int row_in_tile);                  // new work item
//   void tile(uint32_t taddr, int col0, int ncols);                  // (ncols = BN/2 = 64, 80 or 128: policies that
//                                                                    // support BN = 160 mask the 16-column tail of
//                                                                    // their second 64-column chunk)
//                                                                    // this warp's 128-lane x (BN/2)-col
//                                                                    // slice of a finished accumulator:
//                                                                    // taddr = TMEM address of (lane
//                                                                    // quadrant, first column), col0 =
//                                                                    // global n of that column
//   void end(int m_tile, int b, int row_in_tile);                    // work item finished
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace vtm {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UK = 16;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr uint32_t A_BYTES = BM * BK * 2;

template <int BN>
struct Cfg {
  static_assert(BN == 128 || BN == 160 || BN == 256, "BN must be 128, 160 or 256");
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 160 ? 5 : 6);
  // two accumulators of BN columns; the allocation is rounded up to a power of two (320 -> 512 for BN = 160)
  static constexpr uint32_t TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
  static constexpr size_t SMEM_BYTES = 1024 + static_cast<size_t>(STAGES) * STAGE_BYTES + 256;
};

struct Work {
  int batches;          // B
  int m_tiles, n_tiles;
  int tiles_per_split, n_splits;
  int k_chunks;
  int total;
  // Order of the work items.  0: m tile fastest (KA: the dst ranges of one src block run far apart in time, so a later
  // range starts from the maxima the earlier ones published).  1: n split fastest (plain GEMMs): the splits of one m tile
  // run on neighbouring CTAs at the same time, so the A tile comes from HBM once and from L2 afterwards — with m fastest
  // the feed-forward output projection (A = 335 MB > L2) read A from HBM once per n tile.
  int n_fastest;
  __host__ void plan(int M, int N, int K, int B, int BN, int sms, int target_items_per_sm, int min_tiles) {
    batches = B;
    n_fastest = 0;
    m_tiles = (M + BM - 1) / BM;
    n_tiles = (N + BN - 1) / BN;
    k_chunks = (K + BK - 1) / BK;
    const long long base = static_cast<long long>(m_tiles) * B;
    int s = 1;
    while (base * s < static_cast<long long>(target_items_per_sm) * sms && s * 2 <= n_tiles &&
           n_tiles / (s * 2) >= min_tiles)
      s *= 2;
    tiles_per_split = (n_tiles + s - 1) / s;
    n_splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    total = static_cast<int>(base * n_splits);
  }
  __device__ __forceinline__ void decode(int w, int* m_tile, int* b, int* nt0, int* nt1) const {
    int sp;
    if (n_fastest) {
      sp = w % n_splits;
      const int rest = w / n_splits;
      *m_tile = rest % m_tiles;
      *b = rest / m_tiles;
    } else {
      *m_tile = w % m_tiles;
      const int rest = w / m_tiles;
      *b = rest % batches;
      sp = rest / batches;
    }
    *nt0 = sp * tiles_per_split;
    const int e = *nt0 + tiles_per_split;
    *nt1 = e < n_tiles ? e : n_tiles;
  }
};


// ---------------------------------------------------------------------------------------------------------
// Row-contiguous stores for epilogues that write the tile to global memory.  After tcgen05.ld a thread holds ONE
// row of the tile, so a direct 16-byte store per thread touches 32 different cache lines per instruction (ncu:
// L1TEX at 63 % of peak, the busiest unit of the projection GEMMs, and every sector written half at a time).  The
// warp's 32 rows x 64 fp16 columns go through a 4 KB XOR-swizzled staging area instead: each thread writes its row
// as eight 16-byte chunks (chunk k at slot k ^ (row & 7): conflict-free), then lane l reads chunk l & 7 of row
// 4 i + l / 8, so that eight consecutive lanes store 128 contiguous bytes of one row.
constexpr uint32_t STAGE_STORE_BYTES = 32 * 128;
template <class F>
__device__ __forceinline__ void warp_store_rows64(uint32_t scratch, const uint32_t (&pk)[32], F&& store) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(scratch + lane * 128 + ((k ^ (lane & 7)) << 4)),
                 "r"(pk[4 * k]), "r"(pk[4 * k + 1]), "r"(pk[4 * k + 2]), "r"(pk[4 * k + 3])
                 : "memory");
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3), kk = lane & 7;
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(scratch + r * 128 + ((kk ^ (r & 7)) << 4))
                 : "memory");
    store(r, kk * 8, v);   // row within the warp's 32, first of 8 columns within the 64
  }
  __syncwarp();
}

#if defined(VTM_EXP_TRACE)   // timing experiment (tools/ubench/gemm_trace.cu): per-CTA event times in ns
__device__ unsigned long long vtm_trace[160][64];
#define VTM_TRACE(slot)                                                                   \
  do {                                                                                    \
    if ((slot) < 64) {                                                                    \
      unsigned long long t_;                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                              \
      vtm_trace[blockIdx.x][(slot)] = t_;                                                 \
    }                                                                                     \
  } while (0)
#else
#define VTM_TRACE(slot) do { } while (0)
#endif

template <int BN, class Epi>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const Work wk, Epi epi) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128-byte swizzle pattern shared by TMA and UMMA descriptors
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
  // 8-byte slots: full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem base pointer
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) VTM_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
#if defined(VTM_EXP_NO_TMA)
      int exp_chunks = 0;
#endif
      for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
        int m_tile, b, nt0, nt1;
        wk.decode(w, &m_tile, &b, &nt0, &nt1);
        for (int nt = nt0; nt < nt1; ++nt) {
          for (int kc = 0; kc < wk.k_chunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
#if defined(VTM_EXP_NO_TMA)   // timing experiment (tools/ubench/ka_bubbles.cu): only the first ring fill is loaded
            if (exp_chunks++ >= STAGES) { mbar_arrive(full_bar(stage)); } else
#endif
            {
            mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
            tma_load_3d(sa, &tmap_a, full_bar(stage), kc * BK, m_tile * BM, b);
            tma_load_3d(sa + A_BYTES, &tmap_b, full_bar(stage), kc * BK, nt * BN, b);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_ctr = 0;
      for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
        int m_tile, b, nt0, nt1;
        wk.decode(w, &m_tile, &b, &nt0, &nt1);
        for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
          const uint32_t as = tile_ctr & 1u;
          const uint32_t aphase = (tile_ctr >> 1) & 1u;
          mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kc = 0; kc < wk.k_chunks; ++kc) {
#if !defined(VTM_EXP_NO_FULL_WAIT)   // timing experiment: MMA does not wait for the loads
            mbar_wait(full_bar(stage), phase);
#endif
            if (kc == 0) VTM_TRACE(2 + 4 * tile_ctr);          // first operands of the tile have landed
            tc_fence_after();
            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128_kmajor(sa);
            const uint64_t bdesc = umma_desc_sw128_kmajor(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              // +32 bytes per K step inside the swizzle atom = +2 in the (address >> 4) field
              umma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tfull_bar(as));  // accumulator complete
          VTM_TRACE(3 + 4 * tile_ctr);                          // all MMAs of the tile issued
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;     // which BN/2-column half of the tile
    const int row_in_tile = quad * 32 + lane;
    uint32_t tile_ctr = 0;
    Epi e = epi;  // per-thread mutable copy of the policy (kernel parameters are read-only)
    // optional per-warp staging area behind the barriers (epilogues that transpose through shared memory)
    if constexpr (Epi::SCRATCH_PER_WARP > 0)
      e.set_scratch(bar_base + 256u + static_cast<uint32_t>(warp - 2) * Epi::SCRATCH_PER_WARP);
    for (int w = blockIdx.x; w < wk.total; w += gridDim.x) {
      int m_tile, b, nt0, nt1;
      wk.decode(w, &m_tile, &b, &nt0, &nt1);
      e.begin(m_tile, b, row_in_tile);
      for (int nt = nt0; nt < nt1; ++nt, ++tile_ctr) {
        const uint32_t as = tile_ctr & 1u;
        const uint32_t aphase = (tile_ctr >> 1) & 1u;
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        if (warp == 2 && lane == 0) VTM_TRACE(4 + 4 * tile_ctr);   // accumulator complete (MMAs retired)
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * (BN / 2);
#if !defined(VTM_EXP_NO_EPI)   // timing experiment: accumulators are not read
        e.tile(taddr, nt * BN + half * (BN / 2), BN / 2);
#endif
        // every tcgen05.ld issued by tile() has been waited on (tcgen05.wait::ld) before it returns
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
        if (warp == 2 && lane == 0) VTM_TRACE(5 + 4 * tile_ctr);   // epilogue of the tile done (this warp)
      }
      e.end(m_tile, b, row_in_tile);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) VTM_TRACE(1);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, class Epi>
inline int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Work& wk, const Epi& epi, int sms,
                  cudaStream_t stream) {
  using C = Cfg<BN>;
  const size_t smem = C::SMEM_BYTES + static_cast<size_t>(EPI_WARPS) * Epi::SCRATCH_PER_WARP;
  // shared-memory opt-in: once per instantiation and device (a per-function, per-device attribute; immutable afterwards)
  static bool opted[64];
  int dev = 0;
  int rc = cuda_rc(cudaGetDevice(&dev));
  if (rc) return rc;
  if (dev < 0 || dev >= 64 || !opted[dev]) {
    rc = cuda_rc(cudaFuncSetAttribute(gemm_kernel<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    if (rc) return rc;
    if (dev >= 0 && dev < 64) opted[dev] = true;
  }
  const int grid = wk.total < sms ? wk.total : sms;
  gemm_kernel<BN, Epi><<<grid, THREADS, smem, stream>>>(ta, tb, wk, epi);
  return launch_rc();
}

inline int device_sms(int* sms) { return cached_sm_count(sms); }

}  // namespace gemm
}  // namespace vtm
