// Micro-benchmark (r02): issue rates of the instructions a software exp2 needs on sm_100a, per SMSP:
// FFMA, packed fma.rn.f32x2 / add.f32x2 (two fp32 lanes per instruction), HFMA2, IMAD, FMNMX3, MUFU.EX2, and mixes of
// MUFU with the packed polynomial (do they overlap?).  Prints cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float f[16];
  unsigned long long d[8];
  unsigned h[8];
  int n[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) { d[i] = (unsigned long long)__float_as_uint(f[i]) << 32 | __float_as_uint(f[i + 8]); h[i] = 0x3C003C00u + i; n[i] = threadIdx.x + i; }
  const unsigned long long c2 = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (MODE == 0) {  // FFMA 3-reg
#define X(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[15]), "f"(f[14]));
        REP8(X)
#undef X
      }
      if (MODE == 1 || MODE == 6 || MODE == 7) {  // packed FFMA2
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(c2));
        REP8(X)
#undef X
      }
      if (MODE == 2) {  // packed FADD2
#define X(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2));
        REP8(X)
#undef X
      }
      if (MODE == 3) {  // HFMA2
#define X(i) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(0x3C003BFFu));
        REP8(X)
#undef X
      }
      if (MODE == 4) {  // IMAD
#define X(i) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[i]) : "r"(n[7 - i] | 1), "r"(it));
        REP8(X)
#undef X
      }
      if (MODE == 5 || MODE == 6) {  // MUFU.EX2
#define X(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        REP8(X)
#undef X
      }
      if (MODE == 7) {  // MUFU : FFMA2 = 1 : 2 in warp-instr (8 MUFU + 16 FFMA2 per unrolled body): second FFMA2 batch
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(c2));
        REP8(X)
#undef X
#define X(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        REP8(X)
#undef X
      }
      if (MODE == 8) {  // FMNMX3
#define X(i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[15 - i]), "f"(f[(i + 3) & 15]));
        REP8(X)
#undef X
      }
      if (MODE == 10) {  // LOP3 reference for MODE 9's extra xor
#define X(i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(h[i]) : "r"(h[7 - i]));
        REP8(X)
#undef X
      }
      if (MODE == 11 || MODE == 12) {  // 8 MUFU + 8 F2FP(+8 xor): does F2FP share the XU pipe?
#define X(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
        REP8(X)
#undef X
      }
      if (MODE == 9 || MODE == 12) {  // cvt.rn.f16x2.f32 (F2FP) + xor
#define X(i) { unsigned q_; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q_) : "f"(f[i]), "f"(f[i + 8])); h[i] ^= q_; }
        REP8(X)
#undef X
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)(d[i] >> 32) + h[i] + n[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_smsp, int instr_per_body) {
  float* out; long long* cyc;
  int blocks = 148, threads = 128 * warps_per_smsp, iters = 1000;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long hc[148]; cudaMemcpy(hc, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; ++i) c += hc[i]; c /= blocks;
  double per = c / (double(iters) * 4 * instr_per_body * warps_per_smsp);
  printf("%-34s warps/SMSP=%d  cycles per warp-instr per SMSP = %.3f\n", name, warps_per_smsp, per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4}) {
    run<0>("FFMA (3 reg)", w, 8);
    run<1>("FFMA2 fma.rn.f32x2", w, 8);
    run<2>("FADD2 add.rn.f32x2", w, 8);
    run<3>("HFMA2", w, 8);
    run<4>("IMAD", w, 8);
    run<5>("MUFU.EX2", w, 8);
    run<6>("8 FFMA2 + 8 MUFU", w, 16);
    run<7>("16 FFMA2 + 8 MUFU", w, 24);
    run<8>("FMNMX3", w, 8);
    run<9>("8 F2FP + 8 LOP3", w, 16);
    run<10>("LOP3", w, 8);
    run<11>("8 MUFU", w, 8);
    run<12>("8 MUFU + 8 F2FP + 8 LOP3", w, 24);
  }
  return 0;
}
