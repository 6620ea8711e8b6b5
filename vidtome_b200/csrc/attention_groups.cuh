// Flash attention, grouped-softmax variant (r02): ONE CTA per SM with G softmax warp groups.
//
// Why: with head_dim 40 the kernel of attention.cu (2 CTAs/SM x 4 softmax warps) is neither MUFU- nor issue-bound
// but latency-bound — two softmax warps per scheduler cannot cover each other's TMEM-load, barrier and dependency
// stalls (ncu r01: XU 54-72 %, issue 64 %; r02 sweep: moving exponentials to the FMA pipe makes it SLOWER although
// both pipes have room).  Here a CTA owns all 512 TMEM columns and G = 3 softmax groups of 4 warps take the key tiles
// round-robin (tile j -> group j mod G), each with its OWN score buffer, its own un-normalised accumulator O_g and its
// own running reference maximum — an in-CTA split over keys whose partial results are merged once at the end
// (out = sum_g w_g O_g / sum_g w_g l_g with w_g = 2^(m_g - M), the combine of fa_combine_kernel done in TMEM).  A
// scheduler then holds 3 softmax warps that sit in different phases of a tile by construction.
//
//   warp 0       TMA producer (Q once; K/V ring)
//   warp 1       TMEM allocator + single-thread MMA issuer: QK runs G-1 tiles ahead of PV
//   warps 2, 3   idle (complete the first warpgroup so that setmaxnreg can hand its registers to the softmax warps)
//   warps 4+4g.. softmax group g (TMEM lane quadrant = warp % 4)
// TMEM: S_g at [g BKV, (g+1) BKV) (P_g overwrites it as fp16), O_g at [G BKV + g DV_N, ...).
#pragma once

#ifndef VTM_FA_STAGGER_NS
#define VTM_FA_STAGGER_NS 150
#endif

namespace vtm {
namespace {

template <int KSTEPS>
struct FaGroups {
  static constexpr int DV_N = 16 * KSTEPS;
  static constexpr int G = (3 * (64 + DV_N) <= 512) ? 3 : 2;   // head_dim <= 96: three groups
};

template <int KSTEPS>
struct FaGCfg {
  static constexpr int G = FaGroups<KSTEPS>::G;
  static constexpr int ATOMS = (KSTEPS + 3) / 4;
  static constexpr int DV_N = 16 * KSTEPS;
  static constexpr int BKV = (G * (96 + DV_N) <= 512) ? 96 : 64;
  static_assert(G * (BKV + DV_N) <= 512, "TMEM budget");
  static constexpr uint32_t Q_BYTES = ATOMS * BQ * 128;
  static constexpr uint32_t KV_ATOM = BKV * 128;
  static constexpr uint32_t TILE_BYTES = ATOMS * KV_ATOM;
  static constexpr uint32_t STAGE_BYTES = 2 * TILE_BYTES;
  static constexpr int STAGES_RAW = static_cast<int>((196u * 1024u - Q_BYTES) / STAGE_BYTES);
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static_assert(STAGES >= G + 1, "K/V ring must cover the QK run-ahead");
  static constexpr int THREADS = 128 + G * 128;
  static constexpr uint32_t XCH_BYTES = 2 * G * BQ * 4;       // (reference max, denominator) per group and row
  static constexpr size_t SMEM_BYTES = 1024 + Q_BYTES + static_cast<size_t>(STAGES) * STAGE_BYTES + XCH_BYTES + 512;
  static constexpr uint32_t s_col(int g) { return static_cast<uint32_t>(g * BKV); }
  static constexpr uint32_t o_col(int g) { return static_cast<uint32_t>(G * BKV + g * DV_N); }
  // registers: 512 threads leave 128 each; the first warpgroup keeps 40, the softmax warps take 152
  static constexpr bool REALLOC = THREADS == 512;
};

template <int KSTEPS, bool ONES>
__global__ void __launch_bounds__(FaGCfg<KSTEPS>::THREADS, 1)
flash_attn_groups_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                         const __grid_constant__ CUtensorMap tm_v, const FaParams p) {
  using C = FaGCfg<KSTEPS>;
  constexpr int G = C::G;
  constexpr int STAGES = C::STAGES;
  constexpr int BKV = C::BKV;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + C::Q_BYTES;
  const uint32_t xch = sKV + STAGES * C::STAGE_BYTES;          // float xm[G][128], xl[G][128]
  const uint32_t bar_base = xch + C::XCH_BYTES;
  // 8-byte slots: q_full | k_full[S] | v_full[S] | kv_empty[S] | s_full[G] | p_full[G] | o_ready[G] | tmem ptr
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + STAGES + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + 2 * STAGES + s); };
  auto s_full = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + g); };
  auto p_full = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + G + g); };
  auto o_ready = [&](int g) { return bar_base + 8u * (1 + 3 * STAGES + 2 * G + g); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (1 + 3 * STAGES + 3 * G);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int tile = blockIdx.x, part = 0, nparts = 1;
  if (tile >= p.n_full) {
    const int v = tile - p.n_full;
    tile = p.n_full + v / p.splits;
    part = v % p.splits;
    nparts = p.splits;
  }
  const int q0 = (tile % p.n_qtiles) * BQ;
  const int h = (tile / p.n_qtiles) % p.H;
  const int b = tile / (p.n_qtiles * p.H);
  const int bqk = p.qk_shared ? 0 : b;
  const int nkv_all = (p.L + BKV - 1) / BKV;
  const int j_lo = static_cast<int>(static_cast<long long>(part) * nkv_all / nparts);
  const int nkv = static_cast<int>(static_cast<long long>(part + 1) * nkv_all / nparts) - j_lo;   // >= 1

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
    for (int g = 0; g < G; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 4);     // one arrival per softmax warp of the group
      mbar_init(o_ready(g), 1);
    }
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

  if (warp < 4) {
    if (C::REALLOC) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, C::Q_BYTES);
        for (int a = 0; a < C::ATOMS; ++a) tma_load_3d(sQ + a * (BQ * 128), &tm_q, q_full, a * 64, q0, bqk * p.H + h);
      }
      for (int j = 0; j < nkv; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1u;
        mbar_wait(kv_empty(s), ph ^ 1u);
        const uint32_t sk = sKV + s * C::STAGE_BYTES;
        const uint32_t sv = sk + C::TILE_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full(s), C::TILE_BYTES);
          for (int a = 0; a < C::ATOMS; ++a) tma_load_3d(sk + a * C::KV_ATOM, &tm_k, k_full(s), a * 64, (j_lo + j) * BKV, bqk * p.H + h);
          mbar_arrive_expect_tx(v_full(s), C::TILE_BYTES);
          for (int a = 0; a < C::ATOMS; ++a) tma_load_3d(sv + a * C::KV_ATOM, &tm_v, v_full(s), a * 64, (j_lo + j) * BKV, b * p.H + h);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = umma_idesc_f16(BQ, BKV);
      constexpr uint32_t idesc_pv = umma_idesc_f16_bmn(BQ, C::DV_N);
      auto issue_qk = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(k_full(s), (j / STAGES) & 1u);
        tc_fence_after();
        const uint32_t sk = sKV + s * C::STAGE_BYTES;
        const int g = j % G;
        const uint32_t d_tmem = tmem_base + C::s_col(g);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const uint64_t adesc = umma_desc_sw128_kmajor(sQ + (ks >> 2) * (BQ * 128)) + 2u * (ks & 3);
            const uint64_t bdesc = umma_desc_sw128_kmajor(sk + (ks >> 2) * C::KV_ATOM) + 2u * (ks & 3);
            umma_f16(d_tmem, adesc, bdesc, idesc_qk, ks != 0 ? 1u : 0u);
          }
          umma_commit(s_full(g));
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      // S_g of tile j is overwritten by QK(j + G): by then P_j has been consumed by P_j V_j, issued earlier by this
      // thread (tcgen05.mma executes in issue order), so the scores run G-1 tiles ahead without further waits.
      for (int j = 0; j < G - 1 && j < nkv; ++j) issue_qk(j);
      for (int j = 0; j < nkv; ++j) {
        if (j + G - 1 < nkv) issue_qk(j + G - 1);
        const int s = j % STAGES;
        const int g = j % G, it = j / G;
        mbar_wait(v_full(s), (j / STAGES) & 1u);
        mbar_wait(p_full(g), it & 1u);           // P_j in TMEM, O_g rescaled if needed
        tc_fence_after();
        const uint32_t sv = sKV + s * C::STAGE_BYTES + C::TILE_BYTES;
        const uint32_t p_tmem = tmem_base + C::s_col(g);
        const uint32_t o_tmem = tmem_base + C::o_col(g);
        if (elect_one()) {
          const uint64_t vdesc = umma_desc_sw128_mnmajor(sv, C::KV_ATOM);
          if (it == 0) {
#pragma unroll
            for (int t = 0; t < BKV / 16; ++t)
              umma_f16_ts(o_tmem, p_tmem + 8u * t, vdesc + 128u * t, idesc_pv, t != 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int t = 0; t < BKV / 16; ++t)
              umma_f16_ts(o_tmem, p_tmem + 8u * t, vdesc + 128u * t, idesc_pv, 1u);
          }
          umma_commit(kv_empty(s));
          umma_commit(o_ready(g));
        }
        __syncwarp();
      }
    }
  } else {
    if (C::REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    // ===================== softmax group =====================
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int row = q0 + row_in_tile;
    const uint32_t lane_field = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_field + C::s_col(grp);
    const uint32_t o_addr = tmem_base + lane_field + C::o_col(grp);
    const float c = p.scale_log2;
    float m_ref = -INFINITY;
    float l_run = 0.f;
    int it = 0;
#if VTM_FA_STAGGER_NS > 0
    // The groups run identical instruction streams on tiles that become ready together, so left alone they stay in
    // lock-step: all in their MUFU phase at once (sharing the XU pipe), then all in their load / convert / store phase
    // (XU idle).  A one-off offset of a fraction of a tile time per group makes the phases interleave; nothing
    // re-synchronises them afterwards (ncu r02: XU pipe 62 % busy in lock-step).
    if (grp > 0) __nanosleep(VTM_FA_STAGGER_NS * grp);
#endif
    for (int j = grp; j < nkv; j += G, ++it) {
      mbar_wait(s_full(grp), it & 1u);
      tc_fence_after();
      const int n_valid = p.L - (j_lo + j) * BKV;
      const bool tail = n_valid < BKV;
      constexpr int NCH = BKV / 32;
      uint32_t r[NCH][32];
      float alpha = 1.f, sum0 = 0.f, sum1 = 0.f;
      bool redo;
      if (it == 0 || tail) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) tmem_ld_32x32b_x32(s_addr + 32 * ch, r[ch]);
        tmem_ld_wait();
        const float mx = tail ? fa_row_max<NCH, true>(r, n_valid) : fa_row_max<NCH, false>(r, n_valid);
        if ((mx - m_ref) * c > RESCALE_LOG2) {
          alpha = ex2_approx((m_ref - mx) * c);
          m_ref = mx;
        }
        redo = true;
      } else {
        const float mc0 = m_ref * c;
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        tmem_ld_32x32b_x32(s_addr, r[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_wait();
          if (ch + 1 < NCH) tmem_ld_32x32b_x32(s_addr + 32 * (ch + 1), r[ch + 1]);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            m0 = fmax3(m0, __uint_as_float(r[ch][i]), __uint_as_float(r[ch][i + 1]));
            m1 = fmax3(m1, __uint_as_float(r[ch][i + 2]), __uint_as_float(r[ch][i + 3]));
            m2 = fmax3(m2, __uint_as_float(r[ch][i + 4]), __uint_as_float(r[ch][i + 5]));
            m3 = fmax3(m3, __uint_as_float(r[ch][i + 6]), __uint_as_float(r[ch][i + 7]));
          }
          uint32_t pk[16];
          fa_exp32<false, !ONES>(r[ch], pk, c, mc0, 32 * ch, n_valid, sum0, sum1);
          tmem_st_32x32b_x16(s_addr + 16 * ch, pk);
        }
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        const bool moved = (mx - m_ref) * c > RESCALE_LOG2;
        redo = __any_sync(0xffffffffu, moved);
        if (moved) {
          alpha = ex2_approx((m_ref - mx) * c);
          m_ref = mx;
        }
      }
      if (redo) {
        const float mc = m_ref * c;
        sum0 = 0.f;
        sum1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          uint32_t pk[16];
          if (tail) fa_exp32<true, !ONES>(r[ch], pk, c, mc, 32 * ch, n_valid, sum0, sum1);
          else fa_exp32<false, !ONES>(r[ch], pk, c, mc, 32 * ch, n_valid, sum0, sum1);
          tmem_st_32x32b_x16(s_addr + 16 * ch, pk);
        }
      }
      if (!ONES) l_run = l_run * alpha + (sum0 + sum1);
      // O_g rescale: only when a reference max moved.  No wait is needed: the scores of this tile were committed by
      // QK(j), which the MMA thread issued AFTER this group's previous P V (QK runs G-1 tiles ahead, so QK(j) follows
      // PV(j - G) in issue order), and a tcgen05.commit covers every MMA issued before it — having passed s_full for
      // this tile means O_g is at rest.
      if (it > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int cb = 0; cb < C::DV_N; cb += 16) {
          uint32_t ro[16];
          tmem_ld_32x32b_x16(o_addr + cb, ro);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
          tmem_st_32x32b_x16(o_addr + cb, ro);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(grp));
    }
    const int n_my = it;
    // o_ready[g] is waited on exactly once: by now it has completed n_my - 1 or n_my times (the last P V needed this
    // thread's p_full arrival; all earlier ones are covered by the s_full argument above), so the parity of completion
    // n_my - 1 is unambiguous.
    if (n_my >= 1) mbar_wait(o_ready(grp), (n_my - 1) & 1u);
    tc_fence_after();
    // ---- exchange (reference max in log2 units, denominator) and merge the groups' accumulators
    const uint32_t xm = xch + static_cast<uint32_t>((grp * BQ + row_in_tile) * 4);
    const uint32_t xl = xch + static_cast<uint32_t>(((G + grp) * BQ + row_in_tile) * 4);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(xm), "f"(n_my > 0 ? m_ref * c : -INFINITY) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(xl), "f"(l_run) : "memory");
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(G * 128) : "memory");
    tc_fence_after();
    float mg[G], w[G], M = -INFINITY;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mg[g]) : "r"(xch + static_cast<uint32_t>((g * BQ + row_in_tile) * 4)));
      M = fmaxf(M, mg[g]);
    }
    float l = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      w[g] = mg[g] == -INFINITY ? 0.f : ex2_approx(mg[g] - M);
      if (!ONES) {
        float lg;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lg) : "r"(xch + static_cast<uint32_t>(((G + g) * BQ + row_in_tile) * 4)));
        l = fmaf(w[g], lg, l);
      }
    }
    if (ONES) {   // the denominators are column d of the accumulators
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (mg[g] != -INFINITY) {              // warp-uniform per group: a group either had tiles or not
          uint32_t rr[16];
          tmem_ld_32x32b_x16(tmem_base + lane_field + C::o_col(g) + (p.d & ~15), rr);
          tmem_ld_wait();
          l = fmaf(w[g], __uint_as_float((p.d & 15) ? rr[8] : rr[0]), l);   // column d
        }
      }
    }
    const float inv = 1.f / l;
    // column chunks of 16 are dealt round-robin to the groups
    for (int cb = 16 * grp; cb < C::DV_N; cb += 16 * G) {
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (mg[g] != -INFINITY) {
          uint32_t rr[16];
          tmem_ld_32x32b_x16(tmem_base + lane_field + C::o_col(g) + cb, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = fmaf(w[g], __uint_as_float(rr[i]), acc[i]);
        }
      }
      if (nparts == 1) {
        if (row < p.Lq) {
          __half* orow = p.o + (static_cast<size_t>(b) * p.Lq + row) * p.C + h * p.d;
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {
            if (cb + gq * 8 < p.d) {
              uint4 v;
              uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 hh = __floats2half2_rn(acc[gq * 8 + 2 * e] * inv, acc[gq * 8 + 2 * e + 1] * inv);
                pv[e] = *reinterpret_cast<const uint32_t*>(&hh);
              }
              *reinterpret_cast<uint4*>(orow + cb + gq * 8) = v;
            }
          }
        }
      } else {
        const size_t prow = (static_cast<size_t>(tile - p.n_full) * nparts + part) * BQ + row_in_tile;
        float* po = p.part_o + prow * C::DV_N;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq)
          *reinterpret_cast<float4*>(po + cb + 4 * gq) = make_float4(acc[4 * gq], acc[4 * gq + 1], acc[4 * gq + 2], acc[4 * gq + 3]);
        if (cb == 0) p.part_ml[prow] = make_float2(M, l);
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

}  // namespace
}  // namespace vtm
