// KA — fused cosine-similarity GEMM + row arg-max for sm_100a.
//
// Reference semantics (vidtome/merge.py:87,112 / :392,416 and the align_batch variant :93-97):
//   scores = a @ b^T            (fp16 result: fp32 accumulate, round-to-nearest-even to fp16)
//   node_max, node_idx = scores.max(-1)   (first index among equal maxima)
// Here the Ns x Nd score matrix never exists.  The tcgen05 mainloop (gemm_sm100.cuh) gives each CTA a
// 128-row block of src tokens and sweeps dst tiles of 256 tokens; the epilogue below reads the fp32
// accumulators from TMEM (one src row per thread), rounds to fp16 and keeps a running (max, first
// arg).  Partial results are merged across dst ranges / samples with a 64-bit atomicMax on a packed
// (ordered score, ~arg) key, which realises exactly "largest score, then smallest index".
#include <stdlib.h>

#include "gemm_sm100_2cta.cuh"

namespace vtm {
namespace {

constexpr int BN = 256;

struct ArgmaxEpi {
  static constexpr uint32_t SCRATCH_PER_WARP = 0;
  unsigned long long* keys;  // [B', Ns]
  int Ns, Nd, align_batch;
  float best;          // running maximum of this work item, or the entry threshold while best_idx is "none"
  uint32_t best_idx;

  // A work item starts from what other items (other dst ranges / samples of the same src rows) have already
  // published: scores below the published maximum cannot win, so they are filtered with one FMNMX3 per two
  // accumulators and the per-element scan below runs only for the rare 64-column group that holds a candidate.
  // The threshold is the fp16 value just BELOW the published maximum, so an equal score still becomes a candidate
  // and the packed-key atomicMax resolves the tie towards the smaller dst index exactly as before.  Any stale
  // value read here is a valid (lower) threshold: published keys only grow.
  __device__ __forceinline__ void begin(int m_tile, int b, int row_in_tile) {
    best = -INFINITY;
    best_idx = 0xFFFFFFFFu;
    const int row = m_tile * gemm::BM + row_in_tile;
    if (row < Ns) {
      const size_t o = (align_batch ? 0 : static_cast<size_t>(b) * Ns) + row;
      const unsigned long long k = *reinterpret_cast<const volatile unsigned long long*>(keys + o);
      uint32_t ord = static_cast<uint32_t>(k >> 32);
      if (k != 0ull && ord > 0x0400u) {          // published and above -inf
        ord -= 1u;
        if (ord == 0x7FFFu) ord = 0x7FFEu;       // below +0 comes -0, which compares equal: step once more
        best = __half2float(__ushort_as_half(static_cast<unsigned short>(ordered_to_half_bits(ord))));
      }
    }
  }
  static __device__ __forceinline__ float round_f16(float x) { return __half2float(__float2half_rn(x)); }
  __device__ __forceinline__ void scan(const uint32_t (&r)[64], int base, int n_valid) {
#pragma unroll
    for (int c = 0; c < 64; ++c) {
      const float hv = round_f16(__uint_as_float(r[c]));
      if (c < n_valid && hv > best) { best = hv; best_idx = static_cast<uint32_t>(base + c); }
    }
  }
  __device__ __forceinline__ void tile(uint32_t taddr, int col0, int ncols) {
    const int n_valid = Nd - col0;  // columns of this slice that are real dst tokens
#pragma unroll 1
    for (int cb = 0; cb < ncols; cb += 64) {
      uint32_t r[64];
      tmem_ld_32x32b_x64(taddr + cb, r);
      tmem_ld_wait();
      if (cb + 64 <= n_valid) {
        // Rounding to fp16 is monotonic, so max(round(x_i)) == round(max(x_i)): filter on the fp32 maxima of four
        // 16-column groups (8 FMNMX3 each).  The branch is warp-wide, hence the fine groups: a candidate in one
        // lane costs the warp a 16-element scan, not a 64-element one.
        float g[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float m = fmax3(__uint_as_float(r[16 * q]), __uint_as_float(r[16 * q + 1]), __uint_as_float(r[16 * q + 2]));
#pragma unroll
          for (int c = 3; c + 1 < 16; c += 2)
            m = fmax3(m, __uint_as_float(r[16 * q + c]), __uint_as_float(r[16 * q + c + 1]));
          g[q] = fmaxf(m, __uint_as_float(r[16 * q + 15]));
        }
        if (round_f16(fmax3(fmaxf(g[0], g[1]), g[2], g[3])) > best) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (round_f16(g[q]) > best) {      // ascending groups and strict '>' keep the first index among equals
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float hv = round_f16(__uint_as_float(r[16 * q + c]));
                if (hv > best) { best = hv; best_idx = static_cast<uint32_t>(col0 + cb + 16 * q + c); }
              }
            }
          }
        }
      } else if (cb < n_valid) {
        scan(r, col0 + cb, n_valid - cb);
      }
    }
  }
  __device__ __forceinline__ void end(int m_tile, int b, int row_in_tile) {
    const int row = m_tile * gemm::BM + row_in_tile;
    if (row < Ns && best_idx != 0xFFFFFFFFu) {
      const uint32_t hb = static_cast<uint32_t>(__half_as_ushort(__float2half_rn(best)));
      const uint32_t arg =
          align_batch ? static_cast<uint32_t>(b) * static_cast<uint32_t>(Nd) + best_idx : best_idx;
      const size_t o = (align_batch ? 0 : static_cast<size_t>(b) * Ns) + row;
      atomicMax(keys + o, static_cast<unsigned long long>(pack_key(hb, arg)));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Verification twin on CUDA cores: one thread per src row, dst rows staged through shared memory,
// sequential fp32 FMA over K.  O(Ns*Nd*C) scalar work — for tests only.
__global__ void sim_argmax_simt_kernel(const __half* __restrict__ a, const __half* __restrict__ bm, int B,
                                       int Ns, int Nd, int C, int align_batch, unsigned long long* keys) {
  extern __shared__ __half sb[];  // [TJ][C]
  constexpr int TJ = 32;
  const int b = blockIdx.y;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const __half* arow = a + (static_cast<size_t>(b) * Ns + (row < Ns ? row : 0)) * C;
  float best = -INFINITY;
  uint32_t best_idx = 0xFFFFFFFFu;
  for (int j0 = 0; j0 < Nd; j0 += TJ) {
    const int nj = min(TJ, Nd - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < nj * C; e += blockDim.x)
      sb[e] = bm[(static_cast<size_t>(b) * Nd + j0) * C + e];
    __syncthreads();
    if (row < Ns) {
      for (int j = 0; j < nj; ++j) {
        float acc = 0.f;
        for (int k = 0; k < C; ++k) acc = fmaf(__half2float(arow[k]), __half2float(sb[j * C + k]), acc);
        const float hv = __half2float(__float2half_rn(acc));
        if (hv > best) { best = hv; best_idx = j0 + j; }
      }
    }
  }
  if (row < Ns && best_idx != 0xFFFFFFFFu) {
    const uint32_t hb = static_cast<uint32_t>(__half_as_ushort(__float2half_rn(best)));
    const uint32_t arg = align_batch ? static_cast<uint32_t>(b) * Nd + best_idx : best_idx;
    const size_t o = (align_batch ? 0 : static_cast<size_t>(b) * Ns) + row;
    atomicMax(keys + o, static_cast<unsigned long long>(pack_key(hb, arg)));
  }
}

int check_args(const void* a, const void* b, int B, int Ns, int Nd, int C, const void* keys) {
  if (!a || !b || !keys) return VTM_E_NULL;
  if (B <= 0 || Ns <= 0 || Nd <= 0 || C <= 0 || (C % 8) != 0) return VTM_E_SHAPE;
  if (static_cast<long long>(B) * Nd >= 0xFFFFFFFFll) return VTM_E_SHAPE;
  return VTM_OK;
}

}  // namespace
}  // namespace vtm

namespace vtm {
namespace {
int sim_argmax_impl(const void* a_dev, const void* b_dev, int32_t B, int32_t Ns, int32_t Nd, int32_t C,
                    int32_t align_batch, uint64_t* keys_out_dev, void* stream_, bool use_pair) {
  int rc = check_args(a_dev, b_dev, B, Ns, Nd, C, keys_out_dev);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int Bp = align_batch ? 1 : B;
  rc = cuda_rc(cudaMemsetAsync(keys_out_dev, 0, sizeof(uint64_t) * static_cast<size_t>(Bp) * Ns, stream));
  if (rc) return rc;

  CUtensorMap ta, tb;
  rc = make_tmap_3d_f16(&ta, a_dev, C, Ns, B, C, static_cast<uint64_t>(Ns) * C, gemm::BK, gemm::BM);
  if (rc) return rc;
  rc = make_tmap_3d_f16(&tb, b_dev, C, Nd, B, C, static_cast<uint64_t>(Nd) * C, gemm::BK, BN);
  if (rc) return rc;
  int sms = 0;
  rc = gemm::device_sms(&sms);
  if (rc) return rc;

  // Split the dst sweep so that the persistent grid is balanced (>= 16 work items per SM when the
  // problem allows).  The TMA/MMA pipeline runs straight across work-item boundaries, so finer items cost
  // only one 8-byte atomic per src row per item — even single-tile items pay off on the small levels
  // (ds2 level 2: 48 (block, sample) pairs for 148 SMs).
  gemm::Work wk;
  wk.plan(Ns, Nd, C, B, BN, sms, 16, 1);
  ArgmaxEpi epi;
  epi.keys = reinterpret_cast<unsigned long long*>(keys_out_dev);
  epi.Ns = Ns; epi.Nd = Nd; epi.align_batch = align_batch ? 1 : 0;
  epi.best = 0.f; epi.best_idx = 0;
  if (use_pair) {
    // CTA-pair variant: 256-row src blocks per pair, B tensor map with 128-row boxes (each CTA loads half a tile)
    CUtensorMap tb2;
    rc = make_tmap_3d_f16(&tb2, b_dev, C, Nd, B, C, static_cast<uint64_t>(Nd) * C, gemm::BK, 128);
    if (rc) return rc;
    gemm::Work wk2;
    gemm::plan_2cta(&wk2, Ns, Nd, C, B, sms, 16);
    return gemm::launch_2cta<ArgmaxEpi>(ta, tb2, wk2, epi, sms, stream);
  }
  return gemm::launch<BN, ArgmaxEpi>(ta, tb, wk, epi, sms, stream);
}
}  // namespace
}  // namespace vtm

extern "C" int vtm_sim_argmax(const void* a_dev, const void* b_dev, int32_t B, int32_t Ns, int32_t Nd,
                              int32_t C, int32_t align_batch, uint64_t* keys_out_dev, void* stream_) {
  return vtm::sim_argmax_impl(a_dev, b_dev, B, Ns, Nd, C, align_batch, keys_out_dev, stream_, false);
}

extern "C" int vtm_sim_argmax_pair(const void* a_dev, const void* b_dev, int32_t B, int32_t Ns, int32_t Nd,
                                   int32_t C, int32_t align_batch, uint64_t* keys_out_dev, void* stream_) {
  return vtm::sim_argmax_impl(a_dev, b_dev, B, Ns, Nd, C, align_batch, keys_out_dev, stream_, true);
}

extern "C" int vtm_sim_argmax_simt(const void* a_dev, const void* b_dev, int32_t B, int32_t Ns,
                                   int32_t Nd, int32_t C, int32_t align_batch, uint64_t* keys_out_dev,
                                   void* stream_) {
  using namespace vtm;
  int rc = check_args(a_dev, b_dev, B, Ns, Nd, C, keys_out_dev);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int Bp = align_batch ? 1 : B;
  rc = cuda_rc(cudaMemsetAsync(keys_out_dev, 0, sizeof(uint64_t) * static_cast<size_t>(Bp) * Ns, stream));
  if (rc) return rc;
  const int threads = 128;
  dim3 grid((Ns + threads - 1) / threads, B);
  const size_t smem = static_cast<size_t>(32) * C * sizeof(__half);
  if (smem > 48 * 1024) {
    rc = cuda_rc(cudaFuncSetAttribute(sim_argmax_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    if (rc) return rc;
  }
  sim_argmax_simt_kernel<<<grid, threads, smem, stream>>>(
      static_cast<const __half*>(a_dev), static_cast<const __half*>(b_dev), B, Ns, Nd, C,
      align_batch ? 1 : 0, reinterpret_cast<unsigned long long*>(keys_out_dev));
  return launch_rc();
}

extern "C" uint16_t vtm_key_score_half_bits(uint64_t key) {
  return static_cast<uint16_t>(vtm::ordered_to_half_bits(static_cast<uint32_t>(key >> 32)));
}
extern "C" uint32_t vtm_key_arg(uint64_t key) { return 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull); }
