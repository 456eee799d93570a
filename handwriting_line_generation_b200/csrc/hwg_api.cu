// Error plumbing and library-level entry points of the C-ABI.
#include "common.cuh"
#include <string.h>

namespace hwg {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace hwg

extern "C" int hwg_version(void) { return 100; }
extern "C" const char* hwg_last_error(void) { return hwg::g_err; }
extern "C" uint64_t hwg_launch_count(void) { return hwg::g_launches.load(); }
