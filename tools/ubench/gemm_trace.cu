// Per-CTA timeline of the projection GEMM (linear.cu built with -DVTM_EXP_TRACE): when do operands land, when are a
// tile's MMAs issued / retired, when is its epilogue done.  Prints the timeline of a few CTAs and averages, in ns
// relative to the earliest CTA start.
#include <cstdio>
#include <vector>
#include "../../vidtome_b200/csrc/linear.cu"

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 20482, N = argc > 2 ? atoi(argv[2]) : 960, K = argc > 3 ? atoi(argv[3]) : 320;
  void *a, *w, *d;
  cudaMalloc(&a, size_t(M) * K * 2); cudaMalloc(&w, size_t(N) * K * 2); cudaMalloc(&d, size_t(M) * N * 2);
  cudaMemset(a, 0x3c, size_t(M) * K * 2); cudaMemset(w, 0x3c, size_t(N) * K * 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) vtm_linear_f16(a, w, nullptr, M, N, K, d, N, nullptr);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 10; ++r) {
    cudaEventRecord(e0);
    vtm_linear_f16(a, w, nullptr, M, N, K, d, N, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
  }
  printf("M=%d N=%d K=%d: best %.2f us (event pair around one launch)\n", M, N, K, best * 1e3);
  static unsigned long long h[160][64];
  cudaMemcpyFromSymbol(h, vtm::gemm::vtm_trace, sizeof(h));
  unsigned long long t0 = ~0ull, tend = 0;
  for (int c = 0; c < 148; ++c) { if (h[c][0] && h[c][0] < t0) t0 = h[c][0]; if (h[c][1] > tend) tend = h[c][1]; }
  printf("kernel span (first CTA start -> last CTA end): %.2f us\n", (tend - t0) * 1e-3);
  const int show[] = {0, 1, 73, 147};
  for (int c : show) {
    printf("CTA %3d: start %+6lld end %6lld |", c, (long long)(h[c][0] - t0), (long long)(h[c][1] - t0));
    for (int t = 0; t < 10; ++t) {
      if (!h[c][2 + 4 * t]) break;
      printf(" [t%d land %lld issued %lld retired %lld epi %lld]", t, (long long)(h[c][2 + 4 * t] - t0),
             (long long)(h[c][3 + 4 * t] - t0), (long long)(h[c][4 + 4 * t] - t0), (long long)(h[c][5 + 4 * t] - t0));
    }
    printf("\n");
  }
  return 0;
}
