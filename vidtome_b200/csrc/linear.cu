// Projection GEMM of KD: D[M, N] = A[M, K] * W[N, K]^T (+ bias[N]), fp16 in / fp32 accumulate / fp16 out,
// i.e. torch.nn.Linear on fp16 tensors (to_q / to_k / to_v / to_out[0] of the attention module the
// reference calls at vidtome/patch.py:157-162).  Uses the shared tcgen05 mainloop; the epilogue converts
// each thread's accumulator row to fp16 and stores 64-byte runs.
#include "gemm_sm100.cuh"

namespace vtm {
namespace {

struct StoreEpi {
  static constexpr uint32_t SCRATCH_PER_WARP = gemm::STAGE_STORE_BYTES;
  __half* d;
  const __half* bias;
  long long ldd;
  int M, N;
  int row0;            // first of the warp's 32 rows
  uint32_t scratch;

  __device__ __forceinline__ void set_scratch(uint32_t a) { scratch = a; }
  __device__ __forceinline__ void begin(int m_tile, int, int row_in_tile) {
    row0 = m_tile * gemm::BM + (row_in_tile & ~31);
  }
  __device__ __forceinline__ void tile(uint32_t taddr, int col0, int ncols) {
#pragma unroll 1
    for (int cb = 0; cb < ncols; cb += 64) {
      uint32_t r[64];
      tmem_ld_32x32b_x64(taddr + cb, r);
      tmem_ld_wait();
      uint32_t pk[32];
      const int c0 = col0 + cb;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        float f0 = __uint_as_float(r[2 * e]), f1 = __uint_as_float(r[2 * e + 1]);
        if (bias && c0 + 2 * e < N) {          // N % 8 == 0
          f0 += __half2float(bias[c0 + 2 * e]);
          f1 += __half2float(bias[c0 + 2 * e + 1]);
        }
        pk[e] = pack_f16x2(f0, f1);
      }
      // store phase: this lane writes columns [c, c + 8) of rows row0 + lane / 8 + 4 i
      const int c = c0 + (threadIdx.x & 7) * 8;
      int row = row0 + ((threadIdx.x & 31) >> 3);
      __half* dst = d + static_cast<long long>(row) * ldd + c;
      const long long row_step = 4 * ldd;
      gemm::warp_store_rows64(scratch, pk, [&](int, int, const uint4& v) {
        if (row < M && c < N) *reinterpret_cast<uint4*>(dst) = v;
        row += 4;
        dst += row_step;
      });
    }
  }
  __device__ __forceinline__ void end(int, int, int) {}
};

}  // namespace
}  // namespace vtm

extern "C" int vtm_linear_f16(const void* a_dev, const void* w_dev, const void* bias_dev, int32_t M, int32_t N,
                              int32_t K, void* d_dev, int64_t ldd, void* stream_) {
  using namespace vtm;
  if (!a_dev || !w_dev || !d_dev) return VTM_E_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0 || (N % 8) != 0 || (ldd % 8) != 0 || ldd < N)
    return VTM_E_SHAPE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int sms = 0;
  int rc = gemm::device_sms(&sms);
  if (rc) return rc;
  StoreEpi epi;
  epi.d = static_cast<__half*>(d_dev);
  epi.bias = static_cast<const __half*>(bias_dev);
  epi.ldd = ldd; epi.M = M; epi.N = N; epi.row0 = 0; epi.scratch = 0;
  CUtensorMap ta, tb;
  rc = make_tmap_3d_f16(&ta, a_dev, K, M, 1, K, static_cast<uint64_t>(M) * K, gemm::BK, gemm::BM);
  if (rc) return rc;
  gemm::Work wk;
  // BN = 256 unless that leaves a mostly-empty last tile for narrow outputs
  const bool wide = (N % 256 == 0) || N >= 1024;
  if (wide) {
    rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 256);
    if (rc) return rc;
    wk.plan(M, N, K, 1, 256, sms, 16, 1);
    return gemm::launch<256, StoreEpi>(ta, tb, wk, epi, sms, stream);
  }
  rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 128);
  if (rc) return rc;
  wk.plan(M, N, K, 1, 128, sms, 16, 1);
  return gemm::launch<128, StoreEpi>(ta, tb, wk, epi, sms, stream);
}
