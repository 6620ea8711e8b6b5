// KD — merged-token self-attention for sm_100a: QKV projection (tcgen05 GEMM, linear.cu) -> flash
// attention (this file) -> output projection (tcgen05 GEMM).
//
// Reference: `self.attn1(merged_tokens)` at vidtome/patch.py:157-162 = diffusers Attention; the math is
// restated in the reference at utils/pnp_utils.py:47-95: softmax(q k^T * scale) v per head.
//
// Flash-attention kernel, one CTA per (128-row query tile, head, sample), 192 threads, TWO CTAs per SM
// (256 TMEM columns and <= 104 KB of shared memory each) so that one CTA's softmax overlaps the other's MMAs:
//   warp 0      TMA producer: Q tile once, then K/V tiles of BKV keys through a 2-deep ring.  Q, K and V are
//               read straight out of the [B*L, 3C] projection output through rank-4 tensor maps
//               (head_dim, token, head, sample); columns beyond head_dim are zero-filled by TMA.
//   warp 1      MMA issuer (one lane): S_j = Q K_j^T (A, B from shared memory, K-major; only ceil(d/16) K steps),
//               then O += P_j V_j with A = P_j read from TENSOR MEMORY and B = V_j read MN-major from shared
//               memory (no transposed copy of V), plus l += P_j 1 against a constant ones tile, so the softmax
//               denominator is accumulated by the tensor core from the same fp16 P that weights V.
//   warps 2..5  softmax: one query row per thread.  Pass 1 row max; the reference max only moves when the row
//               max exceeds it by more than 2^8 (then O and l are rescaled in TMEM); pass 2 writes
//               P = exp2(s*scale*log2e - m) as fp16 over the consumed score columns using ex2.approx.f16x2
//               (two exponentials per MUFU op).  TMEM loads are software pipelined.  Finally O / l -> fp16.
// TMEM (256 columns): S/P [0, BKV) | O [128, 128 + 16*KSTEPS) | l [128 + 16*KSTEPS, +16).
#include "common.cuh"
#include "ptx.cuh"

extern "C" int vtm_linear_f16(const void*, const void*, const void*, int32_t, int32_t, int32_t, void*, int64_t,
                              void*);

namespace vtm {
namespace {

constexpr int BQ = 128;    // query rows per CTA (UMMA M)
constexpr int FA_THREADS = 192;
constexpr float RESCALE_LOG2 = 8.f;   // move the reference max only when exceeded by 2^8

struct FaParams {
  int L, H, d, C;
  float scale_log2;   // softmax scale * log2(e)
  __half* o;          // [B*L, C]
};

template <int KSTEPS>
struct FaCfg {
  static constexpr int ATOMS = (KSTEPS + 3) / 4;              // 64-wide head_dim blocks
  static constexpr int DV_N = 16 * KSTEPS;                    // UMMA N of P*V (head_dim rounded up to 16)
  static constexpr int BKV = ATOMS == 1 ? 128 : 64;           // keys per tile
  static constexpr int STAGES = 2;
  static constexpr uint32_t Q_BYTES = ATOMS * BQ * 128;       // Q tile: ATOMS x [128 rows x 128 B]
  static constexpr uint32_t KV_ATOM = BKV * 128;              // one 64-wide block of a K or V tile
  static constexpr uint32_t TILE_BYTES = ATOMS * KV_ATOM;
  static constexpr uint32_t STAGE_BYTES = 2 * TILE_BYTES;     // K + V
  static constexpr uint32_t ONES_BYTES = BKV * 128;
  static constexpr uint32_t O_COL = 128;
  static constexpr uint32_t L_COL = 128 + DV_N;
  static constexpr uint32_t TMEM_COLS = (128 + DV_N + 16) <= 256 ? 256 : 512;
  static constexpr size_t SMEM_BYTES = 1024 + Q_BYTES + ONES_BYTES + static_cast<size_t>(STAGES) * STAGE_BYTES + 128;
};

// Row maximum of this thread's BKV scores (TMEM loads pipelined two deep).  TAIL: only the first n_valid exist.
template <int BKV, bool TAIL>
__device__ __forceinline__ float fa_row_max(uint32_t s_addr, int n_valid) {
  float mx = -INFINITY;
  uint32_t ra[32], rb[32];
  tmem_ld_32x32b_x32(s_addr, ra);
#pragma unroll
  for (int cb = 0; cb < BKV; cb += 64) {
    tmem_ld_wait();
    tmem_ld_32x32b_x32(s_addr + cb + 32, rb);
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!TAIL || cb + i < n_valid) mx = fmaxf(mx, __uint_as_float(ra[i]));
    tmem_ld_wait();
    if (cb + 64 < BKV) tmem_ld_32x32b_x32(s_addr + cb + 64, ra);
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!TAIL || cb + 32 + i < n_valid) mx = fmaxf(mx, __uint_as_float(rb[i]));
  }
  return mx;
}

// P = exp2(s*c - mc) as fp16 pairs written over the already consumed score columns (two exps per MUFU op).
template <int BKV, bool TAIL>
__device__ __forceinline__ void fa_write_p(uint32_t s_addr, float c, float mc, int n_valid) {
  uint32_t ra[32], rb[32];
  tmem_ld_32x32b_x32(s_addr, ra);
#pragma unroll
  for (int cb = 0; cb < BKV; cb += 64) {
    uint32_t pk[16];
    tmem_ld_wait();
    tmem_ld_32x32b_x32(s_addr + cb + 32, rb);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float x0 = fmaf(__uint_as_float(ra[2 * i]), c, -mc);
      float x1 = fmaf(__uint_as_float(ra[2 * i + 1]), c, -mc);
      if (TAIL) {
        if (cb + 2 * i >= n_valid) x0 = -INFINITY;
        if (cb + 2 * i + 1 >= n_valid) x1 = -INFINITY;
      }
      pk[i] = ex2_approx_f16x2(pack_f16x2(x0, x1));
    }
    tmem_st_32x32b_x16(s_addr + (cb >> 1), pk);
    tmem_ld_wait();
    if (cb + 64 < BKV) tmem_ld_32x32b_x32(s_addr + cb + 64, ra);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float x0 = fmaf(__uint_as_float(rb[2 * i]), c, -mc);
      float x1 = fmaf(__uint_as_float(rb[2 * i + 1]), c, -mc);
      if (TAIL) {
        if (cb + 32 + 2 * i >= n_valid) x0 = -INFINITY;
        if (cb + 32 + 2 * i + 1 >= n_valid) x1 = -INFINITY;
      }
      pk[i] = ex2_approx_f16x2(pack_f16x2(x0, x1));
    }
    tmem_st_32x32b_x16(s_addr + ((cb + 32) >> 1), pk);
  }
}

template <int KSTEPS>
__global__ void __launch_bounds__(FA_THREADS, 2)
flash_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const FaParams p) {
  using C = FaCfg<KSTEPS>;
  constexpr int STAGES = C::STAGES;
  constexpr int BKV = C::BKV;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sOnes = sQ + C::Q_BYTES;
  const uint32_t sKV = sOnes + C::ONES_BYTES;
  const uint32_t bar_base = sKV + STAGES * C::STAGE_BYTES;
  // 8-byte slots: q_full | k_full[S] | v_full[S] | kv_empty[S] | s_full | p_full | o_done | tmem ptr
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + STAGES + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + 2 * STAGES + s); };
  const uint32_t s_full = bar_base + 8u * (1 + 3 * STAGES);
  const uint32_t p_full = bar_base + 8u * (2 + 3 * STAGES);
  const uint32_t o_done = bar_base + 8u * (3 + 3 * STAGES);
  const uint32_t tmem_ptr_addr = bar_base + 8u * (4 + 3 * STAGES);

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
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);  // one arrival per softmax warp
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, C::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    // ones tile: B operand of the row-sum MMA, MN-major, 128-byte swizzled rows; element 0 of row r = 1.0
    // sits in 16-byte chunk (r % 8) of the row.  Written through the generic proxy, fenced for the MMA below.
    const int t = threadIdx.x - 64;
    for (int i = t; i < BKV * 8; i += 128) {
      const int r = i >> 3, ch = i & 7;
      const uint32_t v0 = (ch == (r & 7)) ? 0x00003C00u : 0u;
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %2, %2};" ::"r"(sOnes + r * 128 + ch * 16), "r"(v0), "r"(0u)
                   : "memory");
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const uint32_t tmem_o = tmem_base + C::O_COL;
  const uint32_t tmem_l = tmem_base + C::L_COL;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
      for (int a = 0; a < C::ATOMS; ++a) tma_load_4d(sQ + a * (BQ * 128), &tm_q, q_full, a * 64, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1u;
        mbar_wait(kv_empty(s), ph ^ 1u);
        const uint32_t sk = sKV + s * C::STAGE_BYTES;
        const uint32_t sv = sk + C::TILE_BYTES;
        mbar_arrive_expect_tx(k_full(s), C::TILE_BYTES);
        for (int a = 0; a < C::ATOMS; ++a) tma_load_4d(sk + a * C::KV_ATOM, &tm_k, k_full(s), a * 64, j * BKV, h, b);
        mbar_arrive_expect_tx(v_full(s), C::TILE_BYTES);
        for (int a = 0; a < C::ATOMS; ++a) tma_load_4d(sv + a * C::KV_ATOM, &tm_v, v_full(s), a * 64, j * BKV, h, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV);
      constexpr uint32_t idesc_pv = umma_idesc_f16_bmn(BQ, C::DV_N);
      constexpr uint32_t idesc_l = umma_idesc_f16_bmn(BQ, 16);
      auto issue_qk = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(k_full(s), (j / STAGES) & 1u);
        tc_fence_after();
        const uint32_t sk = sKV + s * C::STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint64_t adesc = umma_desc_sw128_kmajor(sQ + (ks >> 2) * (BQ * 128)) + 2u * (ks & 3);
          const uint64_t bdesc = umma_desc_sw128_kmajor(sk + (ks >> 2) * C::KV_ATOM) + 2u * (ks & 3);
          umma_f16(tmem_base, adesc, bdesc, idesc_qk, ks != 0 ? 1u : 0u);
        }
        umma_commit(s_full);   // also: every earlier MMA (P_{j-1} V_{j-1}) has retired when this fires
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      const uint64_t onesdesc = umma_desc_sw128_mnmajor(sOnes, 0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % STAGES;
        mbar_wait(p_full, j & 1u);                 // P_j in TMEM, O/l rescaled if needed
        mbar_wait(v_full(s), (j / STAGES) & 1u);
        tc_fence_after();
        const uint32_t sv = sKV + s * C::STAGE_BYTES + C::TILE_BYTES;
        const uint64_t vdesc = umma_desc_sw128_mnmajor(sv, C::KV_ATOM);
#pragma unroll
        for (int t = 0; t < BKV / 16; ++t) {
          const uint32_t acc = (j | t) != 0 ? 1u : 0u;
          umma_f16_ts(tmem_o, tmem_base + 8u * t, vdesc + 128u * t /* 16 rows x 128 B */, idesc_pv, acc);
          umma_f16_ts(tmem_l, tmem_base + 8u * t, onesdesc + 128u * t, idesc_l, acc);
        }
        umma_commit(kv_empty(s));
        // S_{j+1} overwrites the columns P_j lives in: tcgen05.mma executes in issue order, so it cannot
        // pass the P_j V_j products issued just above.
        if (j + 1 < nkv) issue_qk(j + 1);
        else umma_commit(o_done);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const int quad = warp & 3;
    const int row = q0 + quad * 32 + lane;
    const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_field;
    const float c = p.scale_log2;
    float m_ref = -INFINITY;   // reference max (raw score units) used by the exponentials
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1u);
      tc_fence_after();
      const int n_valid = p.L - j * BKV;  // keys of this tile that exist
      const bool tail = n_valid < BKV;
      // ---- pass 1: row maximum
      const float mx = tail ? fa_row_max<BKV, true>(s_addr, n_valid) : fa_row_max<BKV, false>(s_addr, n_valid);
      // ---- reference max moves only when exceeded by 2^RESCALE_LOG2 (P stays <= 256, exact after O/l rescale)
      float alpha = 1.f;
      if ((mx - m_ref) * c > RESCALE_LOG2) {
        alpha = ex2_approx((m_ref - mx) * c);     // 0 on the first tile (m_ref = -inf)
        m_ref = mx;
      }
      const float mc = m_ref * c;
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        // P_{j-1} V_{j-1} has retired (s_full of tile j was committed after it): rescale O and l in place
#pragma unroll
        for (int cb = 0; cb < C::DV_N + 16; cb += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_o + lane_field + cb, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x32b_x16(tmem_o + lane_field + cb, r);
        }
      }
      // ---- pass 2: P over the consumed score columns
      if (tail) fa_write_p<BKV, true>(s_addr, c, mc, n_valid);
      else fa_write_p<BKV, false>(s_addr, c, mc, n_valid);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> fp16 -> o[b, row, h*d : h*d + d]
    mbar_wait(o_done, 0);
    tc_fence_after();
    float inv;
    {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem_l + lane_field, r);
      tmem_ld_wait();
      inv = 1.f / __uint_as_float(r[0]);
    }
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
    tmem_dealloc(tmem_base, C::TMEM_COLS);
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
  rc = make_tmap_4d_f16(&tk, base + C, dims, strides, 64, Cf::BKV);
  if (rc) return rc;
  rc = make_tmap_4d_f16(&tv, base + 2 * C, dims, strides, 64, Cf::BKV);
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
