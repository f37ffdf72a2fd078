// C-ABI bookkeeping: version, build info, error strings, launch counter.
#include <cuda_runtime.h>

#include "common.cuh"

#include <map>
#include <mutex>

namespace re2e {
unsigned long long g_launches = 0;

int ensure_smem(const void *func, size_t bytes) {
  static std::mutex mu;
  static std::map<const void *, size_t> seen;
  if (bytes > 227 * 1024) return RE2E_E_UNSUPPORTED;
  if (bytes <= 48 * 1024) return RE2E_OK;
  std::lock_guard<std::mutex> lk(mu);
  auto it = seen.find(func);
  if (it != seen.end() && it->second >= bytes) return RE2E_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  seen[func] = bytes;
  return RE2E_OK;
}
}  // namespace re2e

extern "C" int re2e_abi_version(void) { return 5; }

extern "C" const char *re2e_build_info(void) {
  return "re2e_b200 sm_100a nvcc " __DATE__ " " __TIME__
#if RE2E_HAS_TCGEN05
         " tcgen05"
#endif
      ;
}

extern "C" const char *re2e_error_string(int code) {
  switch (code) {
    case RE2E_OK: return "ok";
    case RE2E_E_ARG: return "re2e: bad argument (null pointer, non-positive size or misaligned buffer)";
    case RE2E_E_UNSUPPORTED: return "re2e: shape not supported by the sm_100a kernels";
    case RE2E_E_WORKSPACE: return "re2e: workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "re2e: unknown error";
}

extern "C" unsigned long long re2e_launch_count(void) { return re2e::g_launches; }
