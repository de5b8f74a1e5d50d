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
extern "C" int sd_version(void) { return 200; }
extern "C" const char* sd_operand_dtype(void) { return SD_DTYPE_NAME; }
extern "C" int64_t sd_launch_count(void) { return sd::g_launches.load(); }
extern "C" int sd_cuda_available(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0 ? 1 : 0;
}

extern "C" int sd_host_register(void* h_ptr, size_t bytes) {
  SD_REQUIRE(h_ptr && bytes > 0, "sd_host_register: bad argument");
  SD_CUDA_CHECK(cudaHostRegister(h_ptr, bytes, cudaHostRegisterPortable));
  return SD_OK;
}
extern "C" int sd_host_unregister(void* h_ptr) {
  SD_REQUIRE(h_ptr, "sd_host_unregister: null pointer");
  SD_CUDA_CHECK(cudaHostUnregister(h_ptr));
  return SD_OK;
}
extern "C" int sd_copy_d2h_async(void* h_dst, const void* d_src, size_t bytes, void* stream) {
  if (bytes == 0) return SD_OK;
  SD_REQUIRE(h_dst && d_src, "sd_copy_d2h_async: null pointer");
  SD_CUDA_CHECK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return SD_OK;
}
