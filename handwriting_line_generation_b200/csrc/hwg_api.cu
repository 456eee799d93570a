// Error plumbing and library-level entry points of the C-ABI.
#include "common.cuh"
#include <string.h>
#include <mutex>
#include <unordered_set>

namespace hwg {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_last_conv_kernel{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int ensure_max_smem(const void* func) {
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(func)) return HWG_OK;
  int dev = 0, max_optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(max dynamic shared memory): %s", cudaGetErrorString(e));
    return HWG_ERR_CUDA;
  }
  done.insert(func);
  return HWG_OK;
}
}  // namespace hwg

extern "C" int hwg_version(void) { return 100; }
extern "C" const char* hwg_last_error(void) { return hwg::g_err; }
extern "C" uint64_t hwg_launch_count(void) { return hwg::g_launches.load(); }
extern "C" int hwg_last_conv_kernel(void) { return hwg::g_last_conv_kernel.load(); }
