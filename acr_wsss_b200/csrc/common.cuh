// Shared helpers for libacr_b200.so (error reporting, launch checks, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdint>
#include "../../include/acr_b200.h"

namespace acr {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define ACR_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      acr::set_error(__VA_ARGS__);          \
      return (code);                        \
    }                                       \
  } while (0)

#define ACR_CUDA(expr)                                                   \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) {                                             \
      acr::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));    \
      return (int)_e;                                                    \
    }                                                                    \
  } while (0)

// Strided batched fp32 GEMM on CUDA cores (attn_f32.cu): C[m,n] = alpha * sum_k A[m,k] B[k,n].
// Element (z=(b,h), r, c) of a matrix lives at base + b*sb + h*sh + r*sr + c*sc.
struct Mat { long long sb, sh, sr, sc; };
int launch_sgemm(const float* A, Mat la, const float* B, Mat lb, float* C, Mat lc,
                 int M, int N, int K, int batch, int H, float alpha, cudaStream_t st);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Per-kernel device timing for bench.py's roofline (acr_profile_enable / acr_profile_read): while enabled, a KernelTimer
// around a launch records a CUDA event pair on the launching stream.  Off by default and never on during graph capture.
bool profiling_on();
void profile_record(const char* kernel, cudaEvent_t a, cudaEvent_t b);
struct KernelTimer {
  const char* name;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  KernelTimer(const char* n, cudaStream_t s) : name(n), st(s) {
    if (profiling_on() && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, st);
  }
  ~KernelTimer() {
    if (a && b) {
      cudaEventRecord(b, st);
      profile_record(name, a, b);
    }
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` is >= 32 floats of shared memory. Result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

}  // namespace acr
