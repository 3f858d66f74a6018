// Library-level entry points: ABI version, error string, device capability probe.
#include "common.cuh"
#include <cstring>

namespace acr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace acr

extern "C" int acr_abi_version(void) { return ACR_B200_ABI_VERSION; }

extern "C" const char* acr_last_error_string(void) { return acr::g_err; }

extern "C" int acr_device_is_sm100(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}
