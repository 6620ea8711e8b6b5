// KD — merged-token self-attention for sm_100a: QKV projection (tcgen05 GEMM, linear.cu) -> flash
// attention (this file) -> output projection (tcgen05 GEMM).
//
// Reference: `self.attn1(merged_tokens)` at vidtome/patch.py:157-162 = diffusers Attention; the math is
// restated in the reference at utils/pnp_utils.py:47-95: softmax(q k^T * scale) v per head.
//
// Flash-attention kernel, one CTA per (128-row query tile, head, sample), 192 threads:
//   warp 0      TMA producer: Q tile once, then K/V tiles of 128 keys through a STAGES-deep ring.  Q, K and V
//               are read straight out of the [B*L, 3C] projection output through rank-4 tensor maps
//               (head_dim, token, head, sample); columns beyond head_dim are zero-filled by TMA.
//   warp 1      MMA issuer (one lane): S_j = Q K_j^T (A, B from shared memory, K-major) into one of two TMEM
//               score buffers, and O += P_j V_j with A = P_j read from TENSOR MEMORY and B = V_j read
//               MN-major from shared memory (no transposed copy of V is ever made).
//   warps 2..5  softmax: one query row per thread.  Reads S_j from TMEM, keeps the running max / sum in
//               fp32, writes P_j = exp2((s - m) * scale * log2 e) as fp16 over the first half of the same
//               TMEM columns, rescales O in TMEM when a running max moved, and finally normalises and
//               stores O.
// QK_{j+1} is issued before P_j is awaited, so the tensor pipe computes the next scores while the softmax
// warps work.  TMEM: S0 [0,128) | S1 [128,256) | O [256, 256+16*KSTEPS); P_j aliases S_j.
#include "common.cuh"
#include "ptx.cuh"

extern "C" int vtm_linear_f16(const void*, const void*, const void*, int32_t, int32_t, int32_t, void*, int64_t,
                              void*);

namespace vtm {
namespace {

constexpr int BQ = 128;    // query rows per CTA (UMMA M)
constexpr int BKV = 128;   // keys per tile (UMMA N of QK^T, K extent of PV)
constexpr uint32_t ATOM_BYTES = 128 * 128;  // 128 rows x 64 fp16
constexpr int FA_THREADS = 192;

struct FaParams {
  int L, H, d, C;
  float scale_log2;   // softmax scale * log2(e)
  __half* o;          // [B*L, C]
};

template <int KSTEPS>
struct FaCfg {
  static constexpr int ATOMS = (KSTEPS + 3) / 4;              // 64-wide head_dim blocks
  static constexpr int DV_N = 16 * KSTEPS;                    // UMMA N of P*V (head_dim rounded up to 16)
  static constexpr int STAGES = ATOMS == 1 ? 4 : 2;
  static constexpr uint32_t TILE_BYTES = ATOMS * ATOM_BYTES;  // one of Q / K / V tile
  static constexpr uint32_t STAGE_BYTES = 2 * TILE_BYTES;     // K + V
  static constexpr size_t SMEM_BYTES = 1024 + TILE_BYTES + static_cast<size_t>(STAGES) * STAGE_BYTES + 256;
};

template <int KSTEPS>
__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const FaParams p) {
  using C = FaCfg<KSTEPS>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + C::TILE_BYTES;
  const uint32_t bar_base = sKV + STAGES * C::STAGE_BYTES;
  // 8-byte slots: q_full | k_full[S] | v_full[S] | kv_empty[S] | s_full[2] | p_full[2] | o_ready | tmem ptr
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + STAGES + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + 2 * STAGES + s); };
  auto s_full = [&](int s) { return bar_base + 8u * (1 + 3 * STAGES + s); };
  auto p_full = [&](int s) { return bar_base + 8u * (3 + 3 * STAGES + s); };
  const uint32_t o_ready = bar_base + 8u * (5 + 3 * STAGES);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (6 + 3 * STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nkv = (p.L + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 4);  // one arrival per softmax warp
    }
    mbar_init(o_ready, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::TILE_BYTES);
      for (int a = 0; a < C::ATOMS; ++a) tma_load_4d(sQ + a * ATOM_BYTES, &tm_q, q_full, a * 64, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1u;
        mbar_wait(kv_empty(s), ph ^ 1u);
        const uint32_t sk = sKV + s * C::STAGE_BYTES;
        const uint32_t sv = sk + C::TILE_BYTES;
        mbar_arrive_expect_tx(k_full(s), C::TILE_BYTES);
        for (int a = 0; a < C::ATOMS; ++a) tma_load_4d(sk + a * ATOM_BYTES, &tm_k, k_full(s), a * 64, j * BKV, h, b);
        mbar_arrive_expect_tx(v_full(s), C::TILE_BYTES);
        for (int a = 0; a < C::ATOMS; ++a) tma_load_4d(sv + a * ATOM_BYTES, &tm_v, v_full(s), a * 64, j * BKV, h, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV);
      constexpr uint32_t idesc_pv = umma_idesc_f16_bmn(BQ, C::DV_N);
      auto issue_qk = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(k_full(s), (j / STAGES) & 1u);
        tc_fence_after();
        const uint32_t sk = sKV + s * C::STAGE_BYTES;
        const uint32_t d_tmem = tmem_base + (j & 1) * BKV;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint64_t adesc = umma_desc_sw128_kmajor(sQ + (ks >> 2) * ATOM_BYTES) + 2u * (ks & 3);
          const uint64_t bdesc = umma_desc_sw128_kmajor(sk + (ks >> 2) * ATOM_BYTES) + 2u * (ks & 3);
          umma_f16(d_tmem, adesc, bdesc, idesc_qk, ks != 0 ? 1u : 0u);
        }
        umma_commit(s_full(j & 1));
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_qk(j + 1);          // next scores while the softmax warps work on S_j
        const int s = j % STAGES;
        mbar_wait(p_full(j & 1), (j >> 1) & 1u);   // P_j written, O rescaled
        mbar_wait(v_full(s), (j / STAGES) & 1u);
        tc_fence_after();
        const uint32_t sv = sKV + s * C::STAGE_BYTES + C::TILE_BYTES;
        const uint32_t p_tmem = tmem_base + (j & 1) * BKV;   // fp16 P aliases the first 64 columns of S_j
        const uint64_t vdesc = umma_desc_sw128_mnmajor(sv, BKV * 128u);
#pragma unroll
        for (int t = 0; t < BKV / 16; ++t)
          umma_f16_ts(tmem_o, p_tmem + 8u * t, vdesc + 128u * t /* +2048 B */, idesc_pv, (j | t) != 0 ? 1u : 0u);
        umma_commit(kv_empty(s));
        umma_commit(o_ready);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const int quad = warp & 3;
    const int row = q0 + quad * 32 + lane;
    const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;
    const float c = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full(j & 1), (j >> 1) & 1u);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_field + (j & 1) * BKV;
      const int n_valid = p.L - j * BKV;  // keys of this tile that exist
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int cb = 0; cb < BKV; cb += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(s_addr + cb, r);
        tmem_ld_wait();
        if (cb + 32 <= n_valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cb + i < n_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx((m_run - m_new) * c);   // 0 on the first tile (m_run = -inf)
      const float mc = m_new * c;
      // pass 2: P = exp2(s*c - m*c) -> fp16 into TMEM (over the already consumed S columns), row sum
      float sum = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < BKV; cb += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(s_addr + cb, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c, -mc));
          float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c, -mc));
          if (cb + 2 * i >= n_valid) p0 = 0.f;
          if (cb + 2 * i + 1 >= n_valid) p1 = 0.f;
          sum += p0 + p1;
          const __half2 hh = __floats2half2_rn(p0, p1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        tmem_st_32x32b_x16(s_addr + (cb >> 1), pk);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      // rescale O (needs P_{j-1} V_{j-1} retired)
      if (j > 0) {
        mbar_wait(o_ready, (j - 1) & 1u);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int cb = 0; cb < C::DV_N; cb += 16) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem_o + lane_field + cb, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32b_x16(tmem_o + lane_field + cb, r);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(j & 1));
    }
    // epilogue: O / l -> fp16 -> o[b, row, h*d : h*d + d]
    mbar_wait(o_ready, (nkv - 1) & 1u);
    tc_fence_after();
    const float inv = 1.f / l_run;
    __half* orow = p.o + (static_cast<size_t>(b) * p.L + row) * p.C + h * p.d;
#pragma unroll
    for (int cb = 0; cb < C::DV_N; cb += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem_o + lane_field + cb, r);
      tmem_ld_wait();
      if (row < p.L) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (cb + g * 8 < p.d) {  // d % 8 == 0
            uint4 v;
            uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 hh = __floats2half2_rn(__uint_as_float(r[g * 8 + 2 * e]) * inv,
                                                   __uint_as_float(r[g * 8 + 2 * e + 1]) * inv);
              pv[e] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>(orow + cb + g * 8) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int KSTEPS>
int launch_fa(const void* qkv, __half* o, int B, int L, int C, int H, int d, float scale, cudaStream_t stream) {
  using Cf = FaCfg<KSTEPS>;
  CUtensorMap tq, tk, tv;
  const uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(L), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(3) * C, static_cast<uint64_t>(d),
                               static_cast<uint64_t>(L) * 3 * C};
  const __half* base = static_cast<const __half*>(qkv);
  int rc = make_tmap_4d_f16(&tq, base, dims, strides, 64, BQ);
  if (rc) return rc;
  rc = make_tmap_4d_f16(&tk, base + C, dims, strides, 64, BKV);
  if (rc) return rc;
  rc = make_tmap_4d_f16(&tv, base + 2 * C, dims, strides, 64, BKV);
  if (rc) return rc;
  rc = cuda_rc(cudaFuncSetAttribute(flash_attn_kernel<KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(Cf::SMEM_BYTES)));
  if (rc) return rc;
  FaParams p;
  p.L = L; p.H = H; p.d = d; p.C = C;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = o;
  dim3 grid((L + BQ - 1) / BQ, H, B);
  flash_attn_kernel<KSTEPS><<<grid, FA_THREADS, Cf::SMEM_BYTES, stream>>>(tq, tk, tv, p);
  return launch_rc();
}

}  // namespace
}  // namespace vtm

extern "C" size_t vtm_attention_workspace_bytes(int32_t B, int32_t L, int32_t C, int32_t heads) {
  (void)heads;
  if (B <= 0 || L <= 0 || C <= 0) return 0;
  return static_cast<size_t>(B) * L * C * 2 * 4;  // qkv [B*L, 3C] + o [B*L, C], fp16
}

extern "C" int vtm_attention(const void* x_dev, const void* w_qkv_dev, const void* w_o_dev, const void* b_o_dev,
                             int32_t B, int32_t L, int32_t C, int32_t heads, float scale, void* y_dev,
                             void* ws_dev, size_t ws_bytes, void* stream_) {
  using namespace vtm;
  if (!x_dev || !w_qkv_dev || !w_o_dev || !y_dev || !ws_dev) return VTM_E_NULL;
  if (B <= 0 || L <= 0 || C <= 0 || heads <= 0 || C % heads != 0) return VTM_E_SHAPE;
  const int d = C / heads;
  if (d % 8 != 0 || C % 8 != 0) return VTM_E_SHAPE;
  if (d > 128) return VTM_E_UNSUPPORTED;
  if (ws_bytes < vtm_attention_workspace_bytes(B, L, C, heads)) return VTM_E_WS;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  __half* qkv = static_cast<__half*>(ws_dev);
  __half* o = qkv + static_cast<size_t>(B) * L * 3 * C;
  const int M = B * L;
  int rc = vtm_linear_f16(x_dev, w_qkv_dev, nullptr, M, 3 * C, C, qkv, 3 * C, stream_);
  if (rc) return rc;
  const int ksteps = (d + 15) / 16;
  switch (ksteps) {
    case 1: case 2: case 3: rc = launch_fa<3>(qkv, o, B, L, C, heads, d, scale, stream); break;
    case 4: rc = launch_fa<4>(qkv, o, B, L, C, heads, d, scale, stream); break;
    case 5: rc = launch_fa<5>(qkv, o, B, L, C, heads, d, scale, stream); break;
    case 6: rc = launch_fa<6>(qkv, o, B, L, C, heads, d, scale, stream); break;
    default: rc = launch_fa<8>(qkv, o, B, L, C, heads, d, scale, stream); break;
  }
  if (rc) return rc;
  return vtm_linear_f16(o, w_o_dev, b_o_dev, M, C, C, y_dev, C, stream_);
}
