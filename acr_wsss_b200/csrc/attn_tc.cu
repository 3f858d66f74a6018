// placeholder until the tcgen05 kernels land (next commit)
#include "common.cuh"
extern "C" int acr_attn_fwd_bf16(const void*, int, int, int, int, float, void*, float*, float*, long long, float*, void*) {
  acr::set_error("acr_attn_fwd_bf16: not built yet");
  return ACR_E_NOSM100;
}
extern "C" size_t acr_attn_bwd_bf16_workspace(int, int, int, int) { return 256; }
extern "C" int acr_attn_bwd_bf16(const void*, const void*, const float*, const void*, int, int, int, int, float,
                                 const float*, long long, void*, float*, void*, size_t, void*) {
  acr::set_error("acr_attn_bwd_bf16: not built yet");
  return ACR_E_NOSM100;
}
