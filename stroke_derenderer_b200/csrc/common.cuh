// Shared helpers for the sd_b200 library (error plumbing, launch counter).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/sd_b200.h"

namespace sd {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SD_CUDA_CHECK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      sd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SD_ECUDA;                                                               \
    }                                                                                \
  } while (0)

#define SD_LAUNCH_CHECK(name)                                                        \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      sd::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SD_ECUDA;                                                               \
    }                                                                                \
    sd::count_launch();                                                              \
  } while (0)

#define SD_REQUIRE(cond, ...)                                                        \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      sd::set_error(__VA_ARGS__);                                                    \
      return SD_EINVAL;                                                              \
    }                                                                                \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute applies to the current device only: one flag per device for each call site
// (`static PerDeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)`), so that a process driving several
// GPUs raises the shared-memory limit of a kernel on each of them.
struct PerDeviceOnce {
  std::atomic<bool> done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    return !done[dev].exchange(true, std::memory_order_relaxed);
  }
};

}  // namespace sd
