// Process-wide plumbing of the C ABI: error string, version, launch counter.
#include "common.cuh"

namespace sd {
static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace sd

extern "C" const char* sd_last_error(void) { return sd::g_err; }
extern "C" int sd_version(void) { return 100; }
extern "C" int64_t sd_launch_count(void) { return sd::g_launches.load(); }
extern "C" int sd_cuda_available(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0 ? 1 : 0;
}
