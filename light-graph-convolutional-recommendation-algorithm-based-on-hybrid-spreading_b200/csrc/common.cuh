// common.cuh — error plumbing, launch accounting and small device helpers shared by every
// translation unit of liblgcnhs.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lgcnhs.h"

namespace lgc {

// thread-local last-error buffer (lgc_last_error_string)
char* err_buf();
constexpr int kErrBufLen = 512;
void note_launch(int n = 1);

#define LGC_FAIL(code, ...)                                 \
  do {                                                      \
    snprintf(::lgc::err_buf(), ::lgc::kErrBufLen, __VA_ARGS__); \
    return (code);                                          \
  } while (0)

#define LGC_REQUIRE(cond, ...)                       \
  do {                                               \
    if (!(cond)) LGC_FAIL(LGC_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define LGC_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      LGC_FAIL(LGC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,           \
               cudaGetErrorString(_e));                                             \
  } while (0)

// after a kernel launch: catch configuration errors without synchronising
#define LGC_LAUNCH_CHECK(name)                                                      \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess)                                                          \
      LGC_FAIL(LGC_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    ::lgc::note_launch();                                                           \
  } while (0)

// "once per device" guard for cudaFuncSetAttribute: the dynamic-shared-memory opt-in belongs to a device's
// context, so a process that drives several GPUs must set it on each of them.  Thread-safe; a duplicated set from
// two racing threads is harmless.
struct DeviceOnce {
  std::atomic<unsigned long long> done{0ull};
  static int dev() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
  }
  bool need() const { return !((done.load(std::memory_order_acquire) >> dev()) & 1ull); }
  void mark() { done.fetch_or(1ull << dev(), std::memory_order_release); }
};

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace lgc
