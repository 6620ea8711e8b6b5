// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16) as a function of N, for SS (A from smem) and TS
// (A from TMEM) forms, dependent accumulation into one TMEM accumulator (as in a GEMM k-loop / attention PV).
// Operands are garbage (never read back): only timing matters.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vidtome_b200/csrc/ptx.cuh"
using namespace vtm;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, tptr = bar + 8;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tm) : "r"(tptr));
  if (threadIdx.x < 32) {
    const uint32_t idesc = TS ? umma_idesc_f16_bmn(128, N) : umma_idesc_f16(128, N);
    long long t0 = 0, t1 = 0;
    const uint64_t ad = umma_desc_sw128_kmajor(base);
    const uint64_t bd = TS ? umma_desc_sw128_mnmajor(base + 32768, 16384) : umma_desc_sw128_kmajor(base + 32768);
    if (elect_one()) {
      // warm up
      for (int i = 0; i < 8; ++i) { if (TS) umma_f16_ts(tm + 256, tm, bd, idesc, 1); else umma_f16(tm + 256, ad, bd, idesc, 1); }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { if (TS) umma_f16_ts(tm + 256, tm + 8 * (u & 3), bd + 128 * (u & 3), idesc, 1); else umma_f16(tm + 256, ad + 2 * (u & 3), bd + 2 * (u & 3), idesc, 1); }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 1);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
template <int N, bool TS> void run() {
  long long* out; cudaMalloc(&out, 148 * 8);
  const int iters = 2000, smem = 100 * 1024 + 1024;
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TS><<<148, 128, smem>>>(out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%s M=128 N=%3d K=16: %.1f cycles per MMA  (ideal M*N*K/4096 = %.1f)  %s\n", TS ? "TS" : "SS", N, c / (iters * 8.0),
         128.0 * N * 16 / 4096 / 1.0 / 1.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  run<16, false>(); run<48, false>(); run<64, false>(); run<96, false>(); run<128, false>(); run<256, false>();
  run<16, true>(); run<48, true>(); run<64, true>(); run<96, true>(); run<128, true>(); run<256, true>();
  return 0;
}
