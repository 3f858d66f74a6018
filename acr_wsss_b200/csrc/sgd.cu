// Optimiser update of the training step on flat buffers: PolyOptimizer (tool/torchutils.py:10-31) as the reference really
// runs it -- SGD whose momentum slot received the weight-decay value and a poly-decayed learning rate:
//   buf = m * buf + g ;  p += (-lr_t) * buf ;  p16 = bf16(p)      (lr_t read from a device scalar: CUDA-graph friendly)
// One HBM pass (read g, buf, p; write buf, p, p16 = 22 bytes per parameter) instead of four elementwise kernels.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, __nv_bfloat16* __restrict__ p16,
                    long long n4, float momentum, const float* __restrict__ neg_lr) {
  const float nl = __ldg(neg_lr);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 bv = reinterpret_cast<float4*>(buf)[i];
    float4 pv = reinterpret_cast<float4*>(p)[i];
    // same operation order (and no FMA contraction across the two statements) as buf.mul_(m).add_(g); p.addcmul_(buf, -lr)
    bv.x = __fadd_rn(__fmul_rn(bv.x, momentum), gv.x); bv.y = __fadd_rn(__fmul_rn(bv.y, momentum), gv.y);
    bv.z = __fadd_rn(__fmul_rn(bv.z, momentum), gv.z); bv.w = __fadd_rn(__fmul_rn(bv.w, momentum), gv.w);
    pv.x = __fadd_rn(pv.x, __fmul_rn(nl, bv.x)); pv.y = __fadd_rn(pv.y, __fmul_rn(nl, bv.y));
    pv.z = __fadd_rn(pv.z, __fmul_rn(nl, bv.z)); pv.w = __fadd_rn(pv.w, __fmul_rn(nl, bv.w));
    reinterpret_cast<float4*>(buf)[i] = bv;
    reinterpret_cast<float4*>(p)[i] = pv;
    if (p16 != nullptr) {
      __nv_bfloat162 a = __floats2bfloat162_rn(pv.x, pv.y), b = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(p16)[i] = u;
    }
  }
}

}  // namespace

extern "C" int acr_sgd_momentum_step(float* param, const float* grad, float* momentum_buf, void* param_bf16, long long n,
                                     float momentum, const float* neg_lr, void* stream) {
  ACR_REQUIRE(param && grad && momentum_buf && neg_lr, ACR_E_INVAL, "acr_sgd_momentum_step: null pointer");
  ACR_REQUIRE(n > 0 && n % 4 == 0, ACR_E_INVAL, "acr_sgd_momentum_step: n must be a positive multiple of 4");
  ACR_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum_buf) & 15) == 0 && ((uintptr_t)param_bf16 & 7) == 0, ACR_E_ALIGN,
              "acr_sgd_momentum_step: 16-byte aligned fp32 buffers (8-byte bf16) required");
  const long long n4 = n / 4;
  const unsigned grid = (unsigned)std::min<long long>((n4 + 255) / 256, 148LL * 16);
  sgd_momentum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, (__nv_bfloat16*)param_bf16, n4, momentum, neg_lr);
  return acr::check_launch("sgd_momentum_kernel");
}
