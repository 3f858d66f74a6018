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
  static int cached[64] = {0};      // per device: 0 unknown, 1 no, 2 yes (keeps CUDA-graph capture free of attribute queries)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev] == 2;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (dev >= 0 && dev < 64) cached[dev] = (major == 10) ? 2 : 1;
  return major == 10 ? 1 : 0;
}
