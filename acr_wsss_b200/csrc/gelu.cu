// Exact (erf) GELU of the ViT MLP (models/vision_transformer.py: Mlp.act = nn.GELU) on bf16 activations, fp32 math.
//   forward : y = 0.5 x (1 + erf(x / sqrt 2))
//   backward: dx = dy (Phi(x) + x phi(x)), fused with the column sum of dx (= the bias gradient of fc1), so the
//             [M, 4E] gradient is not read a second time.  Partials are folded in a fixed order (deterministic).
// Both kernels are HBM-bound streams: 16-byte loads / stores, 8 elements per thread per step.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr float kRsqrt2 = 0.70710678118654752440f;
constexpr float kRsqrt2Pi = 0.39894228040143267794f;
constexpr int kGeluChunks = 64;       // row chunks of the backward (partial column sums per chunk)

// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) and E = exp(-x^2 / 2) from ONE exponential: Abramowitz-Stegun 7.1.26,
//   erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1 / (1 + p z),  z = |x| / sqrt 2,  |error| <= 1.5e-7
// -- four times below half a bf16 ulp of the results stored here, and 3x fewer instructions than erff() + expf(), which
// made both kernels issue-bound instead of HBM-bound.  The coefficients carry the factor 0.5 of Phi.
__device__ __forceinline__ void phi_and_exp(float x, float& phi, float& e) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * kRsqrt2, ax, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));      // exp(-x^2/2) = 2^(-x^2 log2(e)/2)
  float p = 0.5f * 1.061405429f;
  p = fmaf(p, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float h = p * t * e;                       // 0.5 erfc(|x| / sqrt 2)
  phi = x < 0.f ? h : 1.f - h;
}
__device__ __forceinline__ float gelu_f(float x) {
  float phi, e;
  phi_and_exp(x, phi, e);
  return x * phi;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float phi, e;
  phi_and_exp(x, phi, e);
  return fmaf(x * kRsqrt2Pi, e, phi);
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long nv = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = gelu_f(f[e]);
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {       // ragged tail (n not a multiple of 8)
    const long long i = (nv << 3) + threadIdx.x;
    y[i] = __float2bfloat16(gelu_f(__bfloat162float(x[i])));
  }
}

// CTA = 32 column groups (8 columns each) x 8 row lanes; grid = (ceil(F / 256), row chunks)
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                int M, int F, int rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[8][32][9];
  const int cg = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(M, r0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c < F) {
#pragma unroll 2
    for (int r = r0 + ty; r < r1; r += 8) {
      const size_t off = ((size_t)r * F + c) >> 3;
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x) + off);
      const uint4 gv = __ldg(reinterpret_cast<const uint4*>(dy) + off);
      float xf[8], gf[8];
      unpack8(xv, xf);
      unpack8(gv, gf);
#pragma unroll
      for (int e = 0; e < 8; ++e) gf[e] *= gelu_grad_f(xf[e]);
      const uint4 o = pack8(gf);
      reinterpret_cast<uint4*>(dx)[off] = o;
      // the bias gradient sums the ROUNDED values, i.e. exactly what a separate column sum of dx would see
      unpack8(o, gf);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += gf[e];
    }
  }
  if (partial == nullptr) return;
#pragma unroll
  for (int e = 0; e < 8; ++e) red[ty][cg][e] = acc[e];
  __syncthreads();
  // 256 threads = 256 columns of the block: fold the 8 row lanes in a fixed order
  const int col = threadIdx.x;
  if (blockIdx.x * 256 + col < F) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][col >> 3][col & 7];
    partial[(size_t)blockIdx.y * F + blockIdx.x * 256 + col] = t;
  }
}

__global__ void __launch_bounds__(256)
gelu_colsum_finish_kernel(const float* __restrict__ partial, int nparts, int F, float* __restrict__ out, int accumulate) {
  // 64 columns x 4 part groups per CTA: four times shorter load chains than one thread per column (the kernel is pure
  // latency: 64 partial rows of F floats), folded in a fixed order -> deterministic
  __shared__ float red[4][64];
  const int cl = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  float s0 = 0.f, s1 = 0.f;
  if (c < F) {
    int p = grp;
    for (; p + 4 < nparts; p += 8) {
      s0 += __ldg(partial + (size_t)p * F + c);
      s1 += __ldg(partial + (size_t)(p + 4) * F + c);
    }
    for (; p < nparts; p += 4) s0 += __ldg(partial + (size_t)p * F + c);
  }
  red[grp][cl] = s0 + s1;
  __syncthreads();
  if (grp == 0 && c < F) {
    const float s = (red[0][cl] + red[1][cl]) + (red[2][cl] + red[3][cl]);
    out[c] = accumulate ? out[c] + s : s;
  }
}

}  // namespace

extern "C" int acr_gelu_fwd_bf16(const void* x, void* y, long long n, void* stream) {
  ACR_REQUIRE(x && y, ACR_E_INVAL, "acr_gelu_fwd_bf16: null pointer");
  ACR_REQUIRE(n > 0, ACR_E_INVAL, "acr_gelu_fwd_bf16: n must be positive");
  ACR_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, ACR_E_ALIGN, "acr_gelu_fwd_bf16: 16-byte alignment required");
  const long long nv = (n + 7) >> 3;
  const unsigned grid = (unsigned)std::min<long long>((nv + 255) / 256, 148LL * 32);
  gelu_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
  return acr::check_launch("gelu_fwd_kernel");
}

extern "C" size_t acr_gelu_bwd_workspace(int F) { return F > 0 ? (size_t)kGeluChunks * F * sizeof(float) : 0; }

extern "C" int acr_gelu_bwd_bf16(const void* x, const void* dy, void* dx, int M, int F, float* colsum, int accumulate,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  ACR_REQUIRE(x && dy && dx, ACR_E_INVAL, "acr_gelu_bwd_bf16: null pointer");
  ACR_REQUIRE(M > 0 && F > 0 && F % 8 == 0, ACR_E_INVAL, "acr_gelu_bwd_bf16: F must be a positive multiple of 8");
  ACR_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0, ACR_E_ALIGN, "acr_gelu_bwd_bf16: 16-byte alignment required");
  ACR_REQUIRE(colsum == nullptr || (workspace != nullptr && workspace_bytes >= acr_gelu_bwd_workspace(F)), ACR_E_NOMEM,
              "acr_gelu_bwd_bf16: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = M < kGeluChunks * 8 ? 1 : kGeluChunks;
  const int rows_per_chunk = (M + chunks - 1) / chunks;
  dim3 grid((F + 255) / 256, chunks);
  gelu_bwd_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, M, F, rows_per_chunk,
                                        colsum ? (float*)workspace : nullptr);
  if (int e = acr::check_launch("gelu_bwd_kernel")) return e;
  if (colsum) {
    gelu_colsum_finish_kernel<<<(F + 63) / 64, 256, 0, st>>>((const float*)workspace, chunks, F, colsum, accumulate);
    return acr::check_launch("gelu_colsum_finish_kernel");
  }
  return 0;
}
