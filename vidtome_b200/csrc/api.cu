// Pure-host entry points of the C-ABI (no CUDA calls): version, error text, split arithmetic.
#include <math.h>

#include "common.cuh"

extern "C" int vtm_version(void) { return VTM_VERSION; }

extern "C" const char* vtm_error_string(int code) {
  switch (code) {
    case VTM_OK: return "ok";
    case VTM_E_NULL: return "required pointer is NULL";
    case VTM_E_SHAPE: return "size or divisibility constraint violated";
    case VTM_E_SPLIT: return "inconsistent vtm_split_t";
    case VTM_E_WS: return "workspace too small";
    case VTM_E_DRIVER: return "CUDA driver entry point cuTensorMapEncodeTiled unavailable (no driver?)";
    case VTM_E_UNSUPPORTED: return "configuration not supported by this kernel";
    default: break;
  }
  if (code >= 100000) return "cuTensorMapEncodeTiled failed (CUresult = code - 100000)";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown vidtome_b200 error";
}

extern "C" int vtm_split_counts(const vtm_split_t* split, int32_t* num_src, int32_t* num_dst) {
  vtm::Split sp;
  int rc = vtm::make_split(split, &sp);
  if (rc) return rc;
  if (num_src) *num_src = sp.Ns;
  if (num_dst) *num_dst = sp.Nd;
  return VTM_OK;
}

// merge.py:90  r = min(a.shape[1], int(a.shape[1] * ratio)) — Python float (double) multiply, then
// truncation toward zero.
extern "C" int32_t vtm_merge_count(int32_t num_src, double ratio) {
  const double prod = static_cast<double>(num_src) * ratio;
  long long t = static_cast<long long>(trunc(prod));
  if (t > num_src) t = num_src;
  return static_cast<int32_t>(t);
}
