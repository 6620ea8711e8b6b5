// Micro-benchmark: tcgen05.ld / tcgen05.st throughput (32x32b.x32 = 4 KB per warp instruction) with 4 warps of
// one CTA (one per TMEM lane quadrant) and with 2 CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vidtome_b200/csrc/ptx.cuh"
using namespace vtm;

template <int MODE>   // 0: ld x32 + wait each; 1: 3 ld x32 then one wait; 2: st x16 + wait; 3: ld x16
__global__ void __launch_bounds__(128, 2) k(long long* out, float* sink, int iters) {
  __shared__ uint32_t tptr;
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tptr + ((threadIdx.x >> 5) * 32 << 16);
  uint32_t r[32], acc = 0;
  uint32_t r16[16];
  for (int i = 0; i < 16; ++i) r16[i] = i;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { tmem_ld_32x32b_x32(tm + (i & 1) * 32, r); tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r[q]; }
    if (MODE == 1) {
      uint32_t a[32], b[32];
      tmem_ld_32x32b_x32(tm, r); tmem_ld_32x32b_x32(tm + 32, a); tmem_ld_32x32b_x32(tm + 64, b); tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r[q] ^ a[q] ^ b[q];
    }
    if (MODE == 2) { tmem_st_32x32b_x16(tm + (i & 3) * 16, r16); tmem_st_wait(); }
    if (MODE == 3) { tmem_ld_32x32b_x16(tm + (i & 3) * 16, r16); tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) acc ^= r16[q]; }
  }
  long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + r16[3];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tptr, 256); }
}
template <int MODE> void run(const char* name, int ctas_per_sm, double bytes_per_iter_per_warp) {
  long long* out; float* sink; const int blocks = 148 * ctas_per_sm, iters = 20000;
  cudaMalloc(&out, blocks * 8); cudaMalloc(&sink, blocks * 128 * 4);
  k<MODE><<<blocks, 128>>>(out, sink, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long* h = new long long[blocks]; cudaMemcpy(h, out, blocks * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; ++i) c += h[i]; c /= blocks;
  const double per = c / iters;
  printf("%-34s CTAs/SM=%d: %.1f cycles/iter/warp -> %.0f B/clk/SM %s\n", name, ctas_per_sm, per,
         bytes_per_iter_per_warp * 4 * ctas_per_sm / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  delete[] h; cudaFree(out); cudaFree(sink);
}
int main() {
  for (int c : {1, 2}) {
    run<0>("ld 32x32b.x32 + wait", c, 4096);
    run<1>("3 x ld 32x32b.x32, one wait", c, 3 * 4096);
    run<3>("ld 32x32b.x16 + wait", c, 2048);
    run<2>("st 32x32b.x16 + wait", c, 2048);
  }
  return 0;
}
