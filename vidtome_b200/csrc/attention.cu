// KD — merged-token self-attention (placeholder translation unit: kernel under construction).
#include "common.cuh"

extern "C" size_t vtm_attention_workspace_bytes(int32_t B, int32_t L, int32_t C, int32_t heads) {
  (void)heads;
  if (B <= 0 || L <= 0 || C <= 0) return 0;
  return static_cast<size_t>(B) * L * C * 2 * 4;
}

extern "C" int vtm_attention(const void*, const void*, const void*, const void*, int32_t, int32_t, int32_t,
                             int32_t, float, void*, void*, size_t, void*) {
  return VTM_E_UNSUPPORTED;
}
