// Host-side helpers shared by the newer tcgen05 kernels (conv_gather.cu, conv_head.cu): the driver's tensor-map encoder
// (resolved at run time, so the library loads on a CPU-only box), the SM count, the opt-in for > 48 KB of dynamic shared memory.
#pragma once
#include <cuda.h>

#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace rife {
namespace tch {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// once per (function, device): opt in to `bytes` of dynamic shared memory
inline int opt_in_smem(const void* fn, int bytes) {
  static std::vector<std::pair<const void*, int>> seen;
  static std::mutex mu;   // one host thread per GPU may launch concurrently
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  for (auto& s : seen)
    if (s.first == fn && s.second == dev) return RIFE_OK;
  RIFE_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  seen.emplace_back(fn, dev);
  return RIFE_OK;
}

}  // namespace tch
}  // namespace rife
