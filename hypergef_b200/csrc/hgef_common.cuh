// hgef_common.cuh -- error plumbing, device guard and small helpers shared by the
// translation units of libhgef_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "hgef_b200.h"

namespace hg {

// Records a thread-local message and returns `code` (so: `return set_error(...)`).
int set_error(int code, const char *fmt, ...);

#define HG_CUDA_TRY(expr)                                                                  \
  do {                                                                                     \
    cudaError_t hg_e_ = (expr);                                                            \
    if (hg_e_ != cudaSuccess)                                                              \
      return ::hg::set_error(HG_ECUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(hg_e_), \
                             __FILE__, __LINE__);                                          \
  } while (0)

#define HG_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return ::hg::set_error(HG_EINVAL, __VA_ARGS__); \
  } while (0)

// The reference never sets a device (hgnnaggr_cuda.cu:371-383); every entry point here
// switches to the caller's device for its own duration.
class DeviceGuard {
 public:
  explicit DeviceGuard(int device) {
    ok_ = cudaGetDevice(&prev_) == cudaSuccess;
    if (ok_ && device >= 0 && device != prev_) {
      ok_ = cudaSetDevice(device) == cudaSuccess;
      changed_ = ok_;
    }
  }
  ~DeviceGuard() {
    if (changed_) cudaSetDevice(prev_);
  }
  bool ok() const { return ok_; }

 private:
  int prev_ = 0;
  bool ok_ = false, changed_ = false;
};

// RAII device scratch for one-time (plan / builder) work.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  cudaError_t alloc(size_t n) { return cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)); }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

template <typename T>
constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

}  // namespace hg
