// Projection GEMM of KD: D[M, N] = A[M, K] * W[N, K]^T (+ bias[N]), fp16 in / fp32 accumulate / fp16 out,
// i.e. torch.nn.Linear on fp16 tensors (to_q / to_k / to_v / to_out[0] of the attention module the
// reference calls at vidtome/patch.py:157-162).  Uses the shared tcgen05 mainloop; the epilogue converts
// each thread's accumulator row to fp16 and stores 64-byte runs.
#include "gemm_sm100.cuh"

namespace vtm {
namespace {

struct StoreEpi {
  __half* d;
  const __half* bias;
  long long ldd;
  int M, N;
  int row;

  __device__ __forceinline__ void begin(int m_tile, int, int row_in_tile) {
    row = m_tile * gemm::BM + row_in_tile;
  }
  __device__ __forceinline__ void tile(uint32_t taddr, int col0, int ncols) {
#pragma unroll 1
    for (int cb = 0; cb < ncols; cb += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + cb, r);
      tmem_ld_wait();
      if (row < M) {
        __half* drow = d + static_cast<long long>(row) * ldd;
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // groups of 8 columns = one 16-byte store
          const int c = col0 + cb + g * 8;
          if (c < N) {                 // N % 8 == 0: a group is entirely inside or outside
            uint4 v;
            uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float f0 = __uint_as_float(r[g * 8 + 2 * e]);
              float f1 = __uint_as_float(r[g * 8 + 2 * e + 1]);
              if (bias) {
                f0 += __half2float(bias[c + 2 * e]);
                f1 += __half2float(bias[c + 2 * e + 1]);
              }
              const __half2 h = __floats2half2_rn(f0, f1);
              pv[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(drow + c) = v;
          }
        }
      }
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
  epi.ldd = ldd; epi.M = M; epi.N = N; epi.row = 0;
  CUtensorMap ta, tb;
  rc = make_tmap_3d_f16(&ta, a_dev, K, M, 1, K, static_cast<uint64_t>(M) * K, gemm::BK, gemm::BM);
  if (rc) return rc;
  gemm::Work wk;
  // BN = 256 unless that leaves a mostly-empty last tile for narrow outputs
  const bool wide = (N % 256 == 0) || N >= 1024;
  if (wide) {
    rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 256);
    if (rc) return rc;
    wk.plan(M, N, K, 1, 256, sms, 4, 1);
    return gemm::launch<256, StoreEpi>(ta, tb, wk, epi, sms, stream);
  }
  rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 128);
  if (rc) return rc;
  wk.plan(M, N, K, 1, 128, sms, 4, 1);
  return gemm::launch<128, StoreEpi>(ta, tb, wk, epi, sms, stream);
}
