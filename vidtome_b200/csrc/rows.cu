// HBM-bound row kernels: K0 normalise+split, KC merge gather, KE unmerge gather + residual add.
// All three move whole token rows (C fp16, C % 8 == 0) with 16-byte vector accesses; one warp owns
// one row at a time so that the row-index lookup is done once per row and the accesses of a warp are
// contiguous.  Grids are sized as a multiple of the SM count and stride over rows.
#include "common.cuh"
#include "ptx.cuh"

namespace vtm {
namespace {

constexpr int ROW_THREADS = 256;  // 8 warps per CTA
constexpr int MAX_VEC_PER_LANE = 8;  // supports C up to 32 * 8 * 8 = 2048

__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// ------------------------------------------------------------------ K0: normalise + split
// merge.py:84 `metric = metric / metric.norm(dim=-1, keepdim=True)`; merge.py:76-85 split().
// torch half semantics: norm accumulates in fp32 and is rounded to fp16; the division is carried out
// in fp32 on the fp16-rounded norm and rounded to fp16.
__global__ void __launch_bounds__(ROW_THREADS)
normalize_split_kernel(const __half* __restrict__ x, long long x_bs, const int* __restrict__ rowmap,
                       long long map_bs, Split sp, int B, int C, __half* __restrict__ a_out,
                       __half* __restrict__ b_out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const int vecs = C >> 3;
  const long long rows_per_b = static_cast<long long>(sp.Ns) + sp.Nd;
  const long long total = rows_per_b * B;
  for (long long o = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; o < total;
       o += warps_per_grid) {
    const int b = static_cast<int>(o / rows_per_b);
    const int q = static_cast<int>(o - b * rows_per_b);
    int pos;
    __half* dst;
    if (q < sp.Ns) {
      pos = src_pos(sp, q);
      dst = a_out + (static_cast<long long>(b) * sp.Ns + q) * C;
    } else {
      pos = dst_pos(sp, q - sp.Ns);
      dst = b_out + (static_cast<long long>(b) * sp.Nd + (q - sp.Ns)) * C;
    }
    const int row = rowmap ? rowmap[b * map_bs + pos] : pos;
    const __half* src = x + b * x_bs + static_cast<long long>(row) * C;

    uint4 v[MAX_VEC_PER_LANE];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC_PER_LANE; ++i) {
      const int vi = lane + 32 * i;
      if (vi < vecs) {
        v[i] = ld_nc_16(src + vi * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          ss = fmaf(f.x, f.x, ss);
          ss = fmaf(f.y, f.y, ss);
        }
      }
    }
    ss = warp_sum(ss);
    const float nrm = __half2float(__float2half_rn(sqrtf(ss)));
#pragma unroll
    for (int i = 0; i < MAX_VEC_PER_LANE; ++i) {
      const int vi = lane + 32 * i;
      if (vi < vecs) {
        __half2* h = reinterpret_cast<__half2*>(&v[i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          h[e] = __halves2half2(__float2half_rn(f.x / nrm), __float2half_rn(f.y / nrm));
        }
        st_16(dst + vi * 8, v[i]);
      }
    }
  }
}

// ------------------------------------------------------------------ KC / KE: row gathers
// y[b, i, :] = x[b, map[b, i], :] (+ resid[b, i, :]).  Thread-per-16-bytes, 4 rows in flight.
template <bool ADD>
__global__ void __launch_bounds__(ROW_THREADS)
gather_rows_kernel(const __half* __restrict__ x, long long x_bs, const int* __restrict__ map, long long map_bs,
                   const __half* __restrict__ resid, int B, int L, int C, __half* __restrict__ y,
                   long long y_bs) {
  const int vecs = C >> 3;
  const long long total = static_cast<long long>(B) * L * vecs;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  constexpr int U = 4;
  for (long long t0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t0 < total;
       t0 += stride * U) {
    uint4 v[U], rr[U];
    long long oidx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long t = t0 + u * stride;
      oidx[u] = -1;
      if (t < total) {
        const long long rowg = t / vecs;
        const int vi = static_cast<int>(t - rowg * vecs);
        const int b = static_cast<int>(rowg / L);
        const int i = static_cast<int>(rowg - static_cast<long long>(b) * L);
        const int srow = map ? map[b * map_bs + i] : i;
        v[u] = ld_nc_16(x + b * x_bs + static_cast<long long>(srow) * C + vi * 8);
        oidx[u] = b * y_bs + static_cast<long long>(i) * C + vi * 8;
        if (ADD) rr[u] = ld_nc_16(resid + (static_cast<long long>(b) * L + i) * C + vi * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (oidx[u] >= 0) {
        if (ADD) {
          __half2* a = reinterpret_cast<__half2*>(&v[u]);
          const __half2* r2 = reinterpret_cast<const __half2*>(&rr[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] = __hadd2(a[e], r2[e]);
        }
        st_16(y + oidx[u], v[u]);
      }
    }
  }
}

int grid_for(long long work_items, int per_block, int sms, int blocks_per_sm) {
  long long need = (work_items + per_block - 1) / per_block;
  long long cap = static_cast<long long>(sms) * blocks_per_sm;
  if (need < 1) need = 1;
  if (need > cap) need = cap;  // multiple of the SM count when saturated
  return static_cast<int>(need);
}

int sm_count(int* sms) {
  int dev = 0;
  int rc = cuda_rc(cudaGetDevice(&dev));
  if (rc) return rc;
  return cuda_rc(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
}

}  // namespace
}  // namespace vtm

extern "C" int vtm_normalize_split(const void* x_dev, int64_t x_batch_stride, const int32_t* rowmap_dev,
                                   int64_t rowmap_batch_stride, const vtm_split_t* split, int32_t B,
                                   int32_t C, void* a_out_dev, void* b_out_dev, void* stream_) {
  using namespace vtm;
  Split sp;
  int rc = make_split(split, &sp);
  if (rc) return rc;
  if (!x_dev || (sp.Ns > 0 && !a_out_dev) || (sp.Nd > 0 && !b_out_dev)) return VTM_E_NULL;
  if (B <= 0 || C <= 0 || (C % 8) != 0 || C > 32 * 8 * MAX_VEC_PER_LANE) return VTM_E_SHAPE;
  int sms = 0;
  rc = sm_count(&sms);
  if (rc) return rc;
  const long long rows = (static_cast<long long>(sp.Ns) + sp.Nd) * B;
  const int grid = grid_for(rows, ROW_THREADS / 32, sms, 8);
  normalize_split_kernel<<<grid, ROW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __half*>(x_dev), x_batch_stride, rowmap_dev, rowmap_batch_stride, sp, B, C,
      static_cast<__half*>(a_out_dev), static_cast<__half*>(b_out_dev));
  return launch_rc();
}

extern "C" int vtm_gather_rows(const void* x_dev, int64_t x_batch_stride, const int32_t* map_dev,
                               int64_t map_batch_stride, int32_t B, int32_t L, int32_t C, void* y_dev,
                               int64_t y_batch_stride, void* stream_) {
  using namespace vtm;
  if (!x_dev || !y_dev) return VTM_E_NULL;
  if (B <= 0 || L < 0 || C <= 0 || (C % 8) != 0) return VTM_E_SHAPE;
  if (L == 0) return VTM_OK;
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const long long items = static_cast<long long>(B) * L * (C / 8);
  const int grid = grid_for(items, ROW_THREADS * 4, sms, 8);
  gather_rows_kernel<false><<<grid, ROW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __half*>(x_dev), x_batch_stride, map_dev, map_batch_stride, nullptr, B, L, C,
      static_cast<__half*>(y_dev), y_batch_stride);
  return launch_rc();
}

extern "C" int vtm_unmerge_add(const void* y_dev, int64_t y_batch_stride, const int32_t* map_dev,
                               int64_t map_batch_stride, const void* resid_dev, int32_t B, int32_t N,
                               int32_t C, void* out_dev, void* stream_) {
  using namespace vtm;
  if (!y_dev || !out_dev || !map_dev) return VTM_E_NULL;
  if (B <= 0 || N < 0 || C <= 0 || (C % 8) != 0) return VTM_E_SHAPE;
  if (N == 0) return VTM_OK;
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const long long items = static_cast<long long>(B) * N * (C / 8);
  const int grid = grid_for(items, ROW_THREADS * 4, sms, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const long long out_bs = static_cast<long long>(N) * C;
  if (resid_dev)
    gather_rows_kernel<true><<<grid, ROW_THREADS, 0, st>>>(
        static_cast<const __half*>(y_dev), y_batch_stride, map_dev, map_batch_stride,
        static_cast<const __half*>(resid_dev), B, N, C, static_cast<__half*>(out_dev), out_bs);
  else
    gather_rows_kernel<false><<<grid, ROW_THREADS, 0, st>>>(
        static_cast<const __half*>(y_dev), y_batch_stride, map_dev, map_batch_stride, nullptr, B, N, C,
        static_cast<__half*>(out_dev), out_bs);
  return launch_rc();
}
