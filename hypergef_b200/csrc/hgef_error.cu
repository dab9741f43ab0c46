// hgef_error.cu -- thread-local error message, ABI version, device probe, tuning table.
#include <cstring>
#include <mutex>

#include "hgef_common.cuh"

namespace hg {

// Process-wide tuning overrides (hg_tune_set).  The launchers read their geometry from here instead of
// the environment; an absent name means "built-in default".
namespace {
struct TuneEntry { char name[32]; int value; };
constexpr int kMaxTune = 64;
TuneEntry g_tune[kMaxTune];
int g_ntune = 0;
std::mutex g_tune_mu;
}  // namespace

int tune_get(const char *name, int dflt) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  for (int i = 0; i < g_ntune; ++i)
    if (!strcmp(g_tune[i].name, name)) return g_tune[i].value;
  return dflt;
}

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

int hg_abi_version(void) { return 2; }

int hg_tune_set(const char *name, int32_t value, int32_t clear) {
  if (!name || strlen(name) >= sizeof(hg::g_tune[0].name))
    return hg::set_error(HG_EINVAL, "tune_set: bad name");
  std::lock_guard<std::mutex> lk(hg::g_tune_mu);
  for (int i = 0; i < hg::g_ntune; ++i) {
    if (!strcmp(hg::g_tune[i].name, name)) {
      if (clear) hg::g_tune[i] = hg::g_tune[--hg::g_ntune];
      else hg::g_tune[i].value = value;
      return HG_OK;
    }
  }
  if (clear) return HG_OK;
  if (hg::g_ntune == hg::kMaxTune) return hg::set_error(HG_EINVAL, "tune_set: table full");
  strcpy(hg::g_tune[hg::g_ntune].name, name);
  hg::g_tune[hg::g_ntune++].value = value;
  return HG_OK;
}

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
