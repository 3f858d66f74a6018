// Library-level entry points: ABI version, error string, device capability probe.
#include "common.cuh"
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace acr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {
struct ProfEntry { std::string kernel; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::vector<ProfEntry> g_prof;
bool g_prof_on = false;
}  // namespace
bool profiling_on() { return g_prof_on; }
void profile_record(const char* kernel, cudaEvent_t a, cudaEvent_t b) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back({kernel, a, b});
}
}  // namespace acr

extern "C" void acr_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(acr::g_prof_mu);
  acr::g_prof_on = on != 0;
  if (!on) {
    for (auto& e : acr::g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    acr::g_prof.clear();
  }
}

extern "C" int acr_profile_read(const char* kernel, double* total_ms, long long* launches) {
  ACR_REQUIRE(kernel && total_ms && launches, ACR_E_INVAL, "acr_profile_read: null pointer");
  std::lock_guard<std::mutex> lk(acr::g_prof_mu);
  double tot = 0.0;
  long long n = 0;
  for (auto& e : acr::g_prof) {
    if (e.kernel != kernel) continue;
    ACR_CUDA(cudaEventSynchronize(e.b));
    float ms = 0.f;
    ACR_CUDA(cudaEventElapsedTime(&ms, e.a, e.b));
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *launches = n;
  return 0;
}

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
