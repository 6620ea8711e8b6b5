// Micro-benchmark: does cvt.rn.f16x2.f32 (F2FP) share the MUFU/XU pipe with ex2.approx on sm_100a?
// Prints cycles per warp-instruction per SMSP for MUFU only, F2FP only, and the 2:1 mix used by the softmax.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0 || MODE == 2) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
      }
      if (MODE == 1 || MODE == 2) {
        unsigned p, q;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(a0), "f"(a1));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q) : "f"(a2), "f"(a3));
        acc ^= p ^ q;
      }
      if (MODE == 3) {  // FADD reference
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a0) : "f"(a1));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a2) : "f"(a3));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a1) : "f"(a0));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(a3) : "f"(a2));
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_smsp, int ops_per_iter) {
  float* out; long long* cyc;
  int blocks = 148, threads = 128 * warps_per_smsp, iters = 2000;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; ++i) c += h[i]; c /= blocks;
  double per = c / (double(iters) * 8 * ops_per_iter * warps_per_smsp);
  printf("%-28s warps/SMSP=%d  cycles per warp-instr per SMSP = %.2f\n", name, warps_per_smsp, per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4}) {
    run<0>("MUFU.EX2 only", w, 4);
    run<1>("F2FP only", w, 2);
    run<2>("4 MUFU + 2 F2FP", w, 6);
    run<3>("FADD only", w, 4);
  }
  return 0;
}
