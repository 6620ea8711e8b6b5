// Projection GEMM of KD: D[M, N] = A[M, K] * W[N, K]^T (+ bias[N]), fp16 in / fp32 accumulate / fp16 out,
// i.e. torch.nn.Linear on fp16 tensors (to_q / to_k / to_v / to_out[0] of the attention module the
// reference calls at vidtome/patch.py:157-162).  Uses the shared tcgen05 mainloop; the epilogue converts
// each thread's accumulator row to fp16 and stores 64-byte runs.
#include "gemm_sm100.cuh"
#include "gemm_sm100_2cta.cuh"

namespace vtm {
namespace {

struct StoreEpi {
  static constexpr uint32_t SCRATCH_PER_WARP = gemm::STAGE_STORE_BYTES;
  __half* d;
  const __half* bias;
  const __half* resid;   // optional [M, N] (row stride ldr): D = fp16(fp16(acc + bias) + resid), torch's `lin(x) + h`
  long long ldd, ldr;
  int M, N;
  int row0;            // first of the warp's 32 rows
  uint32_t scratch;

  __device__ __forceinline__ void set_scratch(uint32_t a) { scratch = a; }
  __device__ __forceinline__ void begin(int m_tile, int, int row_in_tile) {
    row0 = m_tile * gemm::BM + (row_in_tile & ~31);
  }
  // As in HeadSplitEpi (attention.cu): with K = 320 .. 1280 this epilogue bounds the GEMM, so whatever is the same for the
  // warp's 32 rows or the chunk's 64 columns is tested once (warp-uniformly) and the common case — a full chunk inside N, all
  // rows inside M — runs without per-store range checks (the first version executed ~630 instructions per warp and chunk).
  __device__ __forceinline__ void tile(uint32_t taddr, int col0, int ncols) {
    const int sub = threadIdx.x & 7;
#pragma unroll 1
    for (int cb = 0; cb < ncols; cb += 64) {
      uint32_t r[64];
      tmem_ld_32x32b_x64(taddr + cb, r);
      tmem_ld_wait();
      uint32_t pk[32];
      const int c0 = col0 + cb;
      const bool full = c0 + 64 <= N && ncols - cb >= 64;       // every column of the chunk exists
      if (bias) {
#pragma unroll
        for (int v8 = 0; v8 < 8; ++v8) {         // eight columns at a time: one 16-byte bias load (N % 8 == 0)
          uint4 bv = make_uint4(0, 0, 0, 0);
          if (full || c0 + 8 * v8 < N) bv = __ldg(reinterpret_cast<const uint4*>(bias + c0 + 8 * v8));
          const __half2* b2 = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 bf = __half22float2(b2[e]);
            float a, b;
            f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(r[8 * v8 + 2 * e]), __uint_as_float(r[8 * v8 + 2 * e + 1])),
                                   f32x2_pack(bf.x, bf.y)), a, b);
            pk[4 * v8 + e] = pack_f16x2(a, b);
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) pk[e] = pack_f16x2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
      }
      // store phase: this lane writes columns [c, c + 8) of rows row0 + lane / 8 + 4 i
      const int c = c0 + sub * 8;
      const bool c_ok = full || (c < N && sub * 8 < ncols - cb);     // BN = 160: the second chunk is 16 columns wide
      int row = row0 + ((threadIdx.x & 31) >> 3);
      __half* dst = d + static_cast<long long>(row) * ldd + c;
      const __half* rs = resid ? resid + static_cast<long long>(row) * ldr + c : nullptr;
      const long long row_step = 4 * ldd, rrow_step = 4 * ldr;
      if (row0 + 31 < M && full) {               // warp-uniform: every row and column of the chunk exists (the staging
                                                 // inside warp_store_rows64 needs the whole warp on the same path)
        if (rs) {
          gemm::warp_store_rows64(scratch, pk, [&](int, int, const uint4& v) {
            uint4 o = v;
            const uint4 rv = *reinterpret_cast<const uint4*>(rs);
            __half2* a = reinterpret_cast<__half2*>(&o);
            const __half2* b2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int e = 0; e < 4; ++e) a[e] = __hadd2(a[e], b2[e]);
            *reinterpret_cast<uint4*>(dst) = o;
            dst += row_step;
            rs += rrow_step;
          });
        } else {
          gemm::warp_store_rows64(scratch, pk, [&](int, int, const uint4& v) {
            *reinterpret_cast<uint4*>(dst) = v;
            dst += row_step;
          });
        }
      } else {
        gemm::warp_store_rows64(scratch, pk, [&](int, int, const uint4& v) {
          if (row < M && c_ok) {
            uint4 o = v;
            if (rs) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rs);
              __half2* a = reinterpret_cast<__half2*>(&o);
              const __half2* b2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int e = 0; e < 4; ++e) a[e] = __hadd2(a[e], b2[e]);
            }
            *reinterpret_cast<uint4*>(dst) = o;
          }
          row += 4;
          dst += row_step;
          if (rs) rs += rrow_step;
        });
      }
    }
  }
  __device__ __forceinline__ void end(int, int, int) {}
};

// gelu(g) = 0.5 g (1 + erf(g / sqrt 2)) with ONE MUFU op: for z = |g| / sqrt 2, log2(erfc(z)) is smooth and close to a
// parabola, so h = erfc(z) / 2 = 2^(p(z) - 1) with a degree-5 polynomial p (p(0) = 0; fitted for minimal error of erf on
// [0, 4.2]: |erf error| <= 6e-7, |gelu error| <= 1.1e-6 — four orders below fp16 resolution), z clamped at 4.2
// (erfc(4.2) = 3e-9).  Then gelu(g) = g (1 - h) for g >= 0 and g h for g < 0.  Evaluated on packed lanes (two values).
__device__ __forceinline__ void gelu_x2(float g0, float g1, float& y0, float& y1) {
  const float z0 = fminf(fabsf(g0) * 0.7071067811865476f, 4.2f), z1 = fminf(fabsf(g1) * 0.7071067811865476f, 4.2f);
  const uint64_t z2 = f32x2_pack(z0, z1);
  uint64_t p2 = f32x2_fma(f32x2_pack(-0.00294415686f, -0.00294415686f), z2, f32x2_pack(0.0295900391f, 0.0295900391f));
  p2 = f32x2_fma(p2, z2, f32x2_pack(-0.148665626f, -0.148665626f));
  p2 = f32x2_fma(p2, z2, f32x2_pack(-0.918509366f, -0.918509366f));
  p2 = f32x2_fma(p2, z2, f32x2_pack(-1.62788901f, -1.62788901f));
  p2 = f32x2_fma(p2, z2, f32x2_pack(-1.f, -1.f));                      // p(z) - 1
  float p0, p1;
  f32x2_unpack(p2, p0, p1);
  const float h0 = ex2_approx(p0), h1 = ex2_approx(p1);                 // erfc(z) / 2
  const float gh0 = g0 * h0, gh1 = g1 * h1;
  y0 = g0 >= 0.f ? g0 - gh0 : gh0;
  y1 = g1 >= 0.f ? g1 - gh1 : gh1;
}

struct GegluEpi {
  static constexpr uint32_t SCRATCH_PER_WARP = gemm::STAGE_STORE_BYTES;
  __half* d;             // [M, N/2]
  const __half* bias;    // [N], interleaved like the weight rows (may be null)
  long long ldd;
  int M, N;              // N = GEMM width (2 x output width), N % 64 == 0
  int row0;
  uint32_t scratch;

  __device__ __forceinline__ void set_scratch(uint32_t a) { scratch = a; }
  __device__ __forceinline__ void begin(int m_tile, int, int row_in_tile) {
    row0 = m_tile * gemm::BM + (row_in_tile & ~31);
  }
  __device__ __forceinline__ void tile(uint32_t taddr, int col0, int ncols) {
    // two 64-column loads ([a32|g32] twice) give 64 outputs = one staged store of 32 rows x 64 columns
#pragma unroll 1
    for (int cb = 0; cb < ncols; cb += 128) {
      uint32_t pk[32];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t r[64];
        tmem_ld_32x32b_x64(taddr + cb + 64 * hlf, r);
        tmem_ld_wait();
        const int c0 = col0 + cb + 64 * hlf;
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {       // eight values + their eight gates at a time: two 16-byte bias loads
          uint4 ba = make_uint4(0, 0, 0, 0), bg = ba;
          if (bias && c0 < N) {
            ba = __ldg(reinterpret_cast<const uint4*>(bias + c0 + 8 * v8));
            bg = __ldg(reinterpret_cast<const uint4*>(bias + c0 + 32 + 8 * v8));
          }
          const __half2* ba2 = reinterpret_cast<const __half2*>(&ba);
          const __half2* bg2 = reinterpret_cast<const __half2*>(&bg);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 8 * v8 + 2 * e;
            const float2 fa = __half22float2(ba2[e]), fg = __half22float2(bg2[e]);
            // fp16 roundings of the unfused pipeline: projection output, gelu output, product.  The adds run on packed
            // fp32 lanes; the product of the two fp16 values is taken by HMUL2 (the fp32 product of two fp16 numbers is
            // exact, so one rounding to fp16 either way).
            float a0, a1, g0, g1;
            f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), f32x2_pack(fa.x, fa.y)), a0, a1);
            f32x2_unpack(f32x2_add(f32x2_pack(__uint_as_float(r[32 + i]), __uint_as_float(r[32 + i + 1])), f32x2_pack(fg.x, fg.y)), g0, g1);
            const uint32_t a16 = pack_f16x2(a0, a1);
            const uint32_t g16 = pack_f16x2(g0, g1);
            const float2 gf = __half22float2(*reinterpret_cast<const __half2*>(&g16));
            float q0, q1;
            gelu_x2(gf.x, gf.y, q0, q1);
            const uint32_t q16 = pack_f16x2(q0, q1);
            const __half2 prod = __hmul2(*reinterpret_cast<const __half2*>(&a16), *reinterpret_cast<const __half2*>(&q16));
            pk[16 * hlf + 4 * v8 + e] = *reinterpret_cast<const uint32_t*>(&prod);
          }
        }
      }
      const int oc0 = (col0 + cb) >> 1;                       // first output column of this 64-wide store
      const int c = oc0 + (threadIdx.x & 7) * 8;
      int row = row0 + ((threadIdx.x & 31) >> 3);
      __half* dst = d + static_cast<long long>(row) * ldd + c;
      const long long row_step = 4 * ldd;
      const int No = N >> 1;
      if (row0 + 31 < M && oc0 + 64 <= No) {     // warp-uniform: the whole 32 x 64 block exists
        gemm::warp_store_rows64(scratch, pk, [&](int, int, const uint4& v) {
          *reinterpret_cast<uint4*>(dst) = v;
          dst += row_step;
        });
      } else {
        gemm::warp_store_rows64(scratch, pk, [&](int, int, const uint4& v) {
          if (row < M && c < No) *reinterpret_cast<uint4*>(dst) = v;
          row += 4;
          dst += row_step;
        });
      }
    }
  }
  __device__ __forceinline__ void end(int, int, int) {}
};

}  // namespace
}  // namespace vtm

#ifndef VTM_LINEAR_PAIR_DEFAULT
#define VTM_LINEAR_PAIR_DEFAULT 0
#endif
namespace vtm {
namespace {
int linear_impl(const void* a_dev, const void* w_dev, const void* bias_dev, const void* resid_dev, long long ldr,
                int M, int N, int K, void* d_dev, long long ldd, cudaStream_t stream) {
  if (!a_dev || !w_dev || !d_dev) return VTM_E_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0 || (N % 8) != 0 || (ldd % 8) != 0 || ldd < N) return VTM_E_SHAPE;
  if (resid_dev && ((ldr % 8) != 0 || ldr < N)) return VTM_E_SHAPE;
  int sms = 0;
  int rc = gemm::device_sms(&sms);
  if (rc) return rc;
  StoreEpi epi;
  epi.d = static_cast<__half*>(d_dev);
  epi.bias = static_cast<const __half*>(bias_dev);
  epi.resid = static_cast<const __half*>(resid_dev);
  epi.ldd = ldd; epi.ldr = ldr; epi.M = M; epi.N = N; epi.row0 = 0; epi.scratch = 0;
  CUtensorMap ta, tb;
  rc = make_tmap_3d_f16(&ta, a_dev, K, M, 1, K, static_cast<uint64_t>(M) * K, gemm::BK, gemm::BM);
  if (rc) return rc;
  gemm::Work wk;
  // Tile width: per k-chunk a CTA moves (128 + BN) rows of operands from L2 for 128 x BN outputs, and these GEMMs are
  // bound by that traffic (K = 320 .. 1280), so the widest tile wins unless it leaves a mostly-empty last tile:
  // N = 320 / 640 (output projections, feed-forward output) split exactly into 160-wide tiles.
  if (N % 160 == 0 && N % 256 != 0 && N < 960) {
    rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 160);
    if (rc) return rc;
    wk.plan(M, N, K, 1, 160, sms, 16, 1);
    wk.n_fastest = 1;
    return gemm::launch<160, StoreEpi>(ta, tb, wk, epi, sms, stream);
  }
  const bool wide = (N % 256 == 0) || N >= 1024;
  const int pair_env = gemm::pair_override();
  if (wide && N % 256 == 0 && M >= 2048 && pair_env != 0 && (pair_env == 1 || VTM_LINEAR_PAIR_DEFAULT)) {
    // CTA-pair mainloop: every B byte fetched once per pair of CTAs (see launch_head_proj in attention.cu)
    rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 128);
    if (rc) return rc;
    gemm::plan_2cta(&wk, M, N, K, 1, sms, 16);
    wk.n_fastest = 1;
    return gemm::launch_2cta<StoreEpi>(ta, tb, wk, epi, sms, stream);
  }
  if (wide) {
    rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 256);
    if (rc) return rc;
    wk.plan(M, N, K, 1, 256, sms, 16, 1);
    wk.n_fastest = 1;
    return gemm::launch<256, StoreEpi>(ta, tb, wk, epi, sms, stream);
  }
  rc = make_tmap_3d_f16(&tb, w_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 128);
  if (rc) return rc;
  wk.plan(M, N, K, 1, 128, sms, 16, 1);
  wk.n_fastest = 1;
  return gemm::launch<128, StoreEpi>(ta, tb, wk, epi, sms, stream);
}
}  // namespace
}  // namespace vtm

extern "C" int vtm_linear_f16(const void* a_dev, const void* w_dev, const void* bias_dev, int32_t M, int32_t N,
                              int32_t K, void* d_dev, int64_t ldd, void* stream_) {
  return vtm::linear_impl(a_dev, w_dev, bias_dev, nullptr, 0, M, N, K, d_dev, ldd, static_cast<cudaStream_t>(stream_));
}

extern "C" int vtm_linear_residual_f16(const void* a_dev, const void* w_dev, const void* bias_dev, const void* resid_dev,
                                       int64_t ldr, int32_t M, int32_t N, int32_t K, void* d_dev, int64_t ldd,
                                       void* stream_) {
  if (!resid_dev) return VTM_E_NULL;
  return vtm::linear_impl(a_dev, w_dev, bias_dev, resid_dev, ldr, M, N, K, d_dev, ldd, static_cast<cudaStream_t>(stream_));
}

extern "C" int vtm_linear_geglu_f16(const void* a_dev, const void* w_il_dev, const void* bias_il_dev, int32_t M,
                                    int32_t N, int32_t K, void* d_dev, int64_t ldd, void* stream_) {
  using namespace vtm;
  if (!a_dev || !w_il_dev || !d_dev) return VTM_E_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0 || (N % 64) != 0 || (ldd % 8) != 0 || ldd < N / 2) return VTM_E_SHAPE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int sms = 0;
  int rc = gemm::device_sms(&sms);
  if (rc) return rc;
  GegluEpi epi;
  epi.d = static_cast<__half*>(d_dev);
  epi.bias = static_cast<const __half*>(bias_il_dev);
  epi.ldd = ldd; epi.M = M; epi.N = N; epi.row0 = 0; epi.scratch = 0;
  CUtensorMap ta, tb;
  rc = make_tmap_3d_f16(&ta, a_dev, K, M, 1, K, static_cast<uint64_t>(M) * K, gemm::BK, gemm::BM);
  if (rc) return rc;
  gemm::Work wk;
  const int pair_env = gemm::pair_override();
  if (N % 256 == 0 && M >= 2048 && pair_env != 0 && (pair_env == 1 || VTM_LINEAR_PAIR_DEFAULT)) {
    rc = make_tmap_3d_f16(&tb, w_il_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 128);
    if (rc) return rc;
    gemm::plan_2cta(&wk, M, N, K, 1, sms, 16);
    wk.n_fastest = 1;
    return gemm::launch_2cta<GegluEpi>(ta, tb, wk, epi, sms, stream);
  }
  rc = make_tmap_3d_f16(&tb, w_il_dev, K, N, 1, K, static_cast<uint64_t>(N) * K, gemm::BK, 256);
  if (rc) return rc;
  wk.plan(M, N, K, 1, 256, sms, 16, 1);
  wk.n_fastest = 1;
  return gemm::launch<256, GegluEpi>(ta, tb, wk, epi, sms, stream);
}
