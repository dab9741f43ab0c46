// hgef_error.cu -- thread-local error message, ABI version, device probe.
#include "hgef_common.cuh"

namespace hg {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace hg

extern "C" {

const char *hg_last_error(void) { return hg::g_err; }

int hg_abi_version(void) { return 1; }

int hg_device_cc(int device) {
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) {
    cudaGetLastError();
    hg::set_error(HG_ECUDA, "hg_device_cc: no usable CUDA device %d", device);
    return -1;
  }
  return major * 10 + minor;
}

}  // extern "C"
