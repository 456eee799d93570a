// Shared helpers for libhwg_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/hwg_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhwg_b200 is written for sm_100a only"
#endif

namespace hwg {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_last_conv_kernel;   // 1 = tcgen05 implicit GEMM, 2 = staged-tile mma.sync kernel

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HWG_ERR_CUDA;
  }
  return HWG_OK;
}

#define HWG_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      hwg::set_error(__VA_ARGS__);        \
      return HWG_ERR_INVALID;             \
    }                                     \
  } while (0)

// Dynamic shared memory opt-in, done once per kernel (to the device maximum) so that steady-state launches make
// no attribute calls — required for the launches to be capturable in a CUDA graph.
int ensure_max_smem(const void* func);
#define HWG_SMEM_OPTIN(kernel)                                               \
  do {                                                                       \
    int rc__ = hwg::ensure_max_smem(reinterpret_cast<const void*>(kernel));  \
    if (rc__) return rc__;                                                   \
  } while (0)

#define HWG_CUDA(call)                                                  \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) {                                           \
      hwg::set_error("%s: %s", #call, cudaGetErrorString(e__));         \
      return HWG_ERR_CUDA;                                              \
    }                                                                   \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace hwg
