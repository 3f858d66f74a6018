// LayerNorm forward / backward around the attention block (the glue either side of a1).
// Reference: the nn.LayerNorm(eps=1e-6) calls of Block.forward, models/vision_transformer.py:230-233 (:299 for eps).
//
// HBM-bound.  Forward reads the fp32 residual stream once and writes the normalised rows directly in the GEMM's
// operand type (bf16 on the fused path -- no separate cast kernel), plus mean / rstd per row.  Backward is a single
// pass over (dy, x): one warp per row computes dx with warp-shuffle row reductions while every lane accumulates the
// d-gamma / d-beta partial sums of the columns it owns in registers; CTAs write one partial row each and a second tiny
// kernel folds them in a fixed order (deterministic, no atomics).  Stock PyTorch spends 105 us per call in its
// gamma/beta backward at [12560, 768]; the whole backward here is bounded by ~130 MB of traffic.
#include "attn_tc.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int kWarpsPerCta = 8;

// four consecutive elements of a row, fp32 or bf16 storage
template <bool BF16>
__device__ __forceinline__ float4 load4(const void* __restrict__ base, long long idx4) {
  if (BF16) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(base) + idx4);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return __ldg(reinterpret_cast<const float4*>(base) + idx4);
}
template <bool BF16>
__device__ __forceinline__ void store4(void* __restrict__ base, long long idx4, float4 v) {
  if (BF16) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    reinterpret_cast<uint2*>(base)[idx4] = u;
  } else {
    reinterpret_cast<float4*>(base)[idx4] = v;
  }
}

template <int VEC, bool OUT_BF16, bool X_BF16>      // VEC float4 per lane: E = 128 * VEC
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ res, void* __restrict__ sum_out, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int M, float eps, void* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  constexpr int E = 128 * VEC;
  const int lane = threadIdx.x & 31;
  // grid-stride over rows: the grid is sized to the resident warps of the device (no ragged last wave)
  for (long long row = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * kWarpsPerCta) {
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = load4<X_BF16>(x, row * (E / 4) + i * 32 + lane);
    if (res != nullptr) {       // residual add fused in: s = x + res is stored (rounded to the stream's type) and normalised
      const float4 r = load4<X_BF16>(res, row * (E / 4) + i * 32 + lane);
      v[i] = make_float4(v[i].x + r.x, v[i].y + r.y, v[i].z + r.z, v[i].w + r.w);
      if (X_BF16) {
        v[i].x = __bfloat162float(__float2bfloat16(v[i].x)); v[i].y = __bfloat162float(__float2bfloat16(v[i].y));
        v[i].z = __bfloat162float(__float2bfloat16(v[i].z)); v[i].w = __bfloat162float(__float2bfloat16(v[i].w));
      }
      store4<X_BF16>(sum_out, row * (E / 4) + i * 32 + lane, v[i]);
    }
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = acr::warp_sum(s) * (1.f / E);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rs = rsqrtf(acr::warp_sum(ss) * (1.f / E) + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    const float o0 = (v[i].x - mu) * rs * g.x + bb.x, o1 = (v[i].y - mu) * rs * g.y + bb.y;
    const float o2 = (v[i].z - mu) * rs * g.z + bb.z, o3 = (v[i].w - mu) * rs * g.w + bb.w;
    store4<OUT_BF16>(y, row * (E / 4) + i * 32 + lane, make_float4(o0, o1, o2, o3));
  }
  }
}

template <int VEC, bool DY_BF16, bool X_BF16, bool COLSUM = false>      // X_BF16: x is bf16 and dx is written as bf16
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ dres, const void* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, int M, int rows_per_cta,
                     void* __restrict__ dx, float* __restrict__ part_g, float* __restrict__ part_b, float* __restrict__ part_c) {
  constexpr int E = 128 * VEC;
  constexpr int NARR = COLSUM ? 3 : 2;
  // Column-owned partial sums (d-gamma, d-beta, column sums of dx) live in SHARED memory, one private strip per warp
  // (float4 per lane -> conflict-free, no atomics): in registers they cost 72 registers per thread at E = 768 and held
  // the kernel to 8 warps per SM; this way 2 CTAs (16 warps) are resident and the loads of more rows are in flight.
  extern __shared__ float4 acc_sm[];          // [NARR][kWarpsPerCta][VEC * 32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* ag = acc_sm + (0 * kWarpsPerCta + warp) * (VEC * 32) + lane;
  float4* ab = acc_sm + (1 * kWarpsPerCta + warp) * (VEC * 32) + lane;
  float4* ac = acc_sm + ((COLSUM ? 2 : 1) * kWarpsPerCta + warp) * (VEC * 32) + lane;
  float4 g[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    ag[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (COLSUM) ac[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min((long long)M, r0 + rows_per_cta);
  for (long long row = r0 + warp; row < r1; row += kWarpsPerCta) {
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float4 d[VEC], xh[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      d[i] = load4<DY_BF16>(dy, row * (E / 4) + i * 32 + lane);
      const float4 xv = load4<X_BF16>(x, row * (E / 4) + i * 32 + lane);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 tb = ab[i * 32], tg = ag[i * 32];
      tb.x += d[i].x; tb.y += d[i].y; tb.z += d[i].z; tb.w += d[i].w;
      tg.x += d[i].x * xh[i].x; tg.y += d[i].y * xh[i].y; tg.z += d[i].z * xh[i].z; tg.w += d[i].w * xh[i].w;
      ab[i * 32] = tb;
      ag[i * 32] = tg;
      d[i].x *= g[i].x; d[i].y *= g[i].y; d[i].z *= g[i].z; d[i].w *= g[i].w;       // dy * gamma
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
    }
    s1 = acr::warp_sum(s1) * (1.f / E);
    s2 = acr::warp_sum(s2) * (1.f / E);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float4 o;
      o.x = rs * (d[i].x - s1 - xh[i].x * s2); o.y = rs * (d[i].y - s1 - xh[i].y * s2);
      o.z = rs * (d[i].z - s1 - xh[i].z * s2); o.w = rs * (d[i].w - s1 - xh[i].w * s2);
      if (dres != nullptr) {    // gradient arriving over the residual connection, added before the single rounding
        const float4 r = load4<X_BF16>(dres, row * (E / 4) + i * 32 + lane);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      store4<X_BF16>(dx, row * (E / 4) + i * 32 + lane, o);
      if (COLSUM) {                  // sums what was stored (rounded when the stream is bf16), like a separate column sum would
        if (X_BF16) {
          o.x = __bfloat162float(__float2bfloat16(o.x)); o.y = __bfloat162float(__float2bfloat16(o.y));
          o.z = __bfloat162float(__float2bfloat16(o.z)); o.w = __bfloat162float(__float2bfloat16(o.w));
        }
        float4 tc_ = ac[i * 32];
        tc_.x += o.x; tc_.y += o.y; tc_.z += o.z; tc_.w += o.w;
        ac[i * 32] = tc_;
      }
    }
  }
  // fold the CTA's warps in a fixed order: one partial row per CTA and array
  __syncthreads();
  const float* accf = reinterpret_cast<const float*>(acc_sm);
  for (int a = 0; a < NARR; ++a) {
    float* dst = (a == 0 ? part_g : (a == 1 ? part_b : part_c)) + (long long)blockIdx.x * E;
    for (int c = threadIdx.x; c < E; c += kWarpsPerCta * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerCta; ++w) t += accf[(a * kWarpsPerCta + w) * E + c];
      dst[c] = t;
    }
  }
}

// 8 columns per CTA (one 32-byte sector per partial row), 128 groups of partial rows per column, folded through shared
// memory in a fixed order -> deterministic.  (E / 8 CTAs instead of E / 32: the kernel is latency bound, 24 CTAs left most
// of the machine idle for 8.5 us per LayerNorm.)
__global__ void __launch_bounds__(1024)
layernorm_bwd_finish_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b, const float* __restrict__ part_c, int nparts,
                            int E, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx_colsum, int accumulate) {
  __shared__ float sg[128][9], sb[128][9], sc[128][9];
  const int cl = threadIdx.x & 7, grp = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  float ag = 0.f, ab = 0.f, ac = 0.f;
  if (c < E) {
#pragma unroll 3
    for (int p = grp; p < nparts; p += 128) {
      ag += __ldg(part_g + (long long)p * E + c);
      ab += __ldg(part_b + (long long)p * E + c);
      if (part_c != nullptr) ac += __ldg(part_c + (long long)p * E + c);
    }
  }
  sg[grp][cl] = ag;
  sb[grp][cl] = ab;
  sc[grp][cl] = ac;
  __syncthreads();
  // 24 threads finish: (array, column); 128 values each, 4 independent chains
  if (threadIdx.x < 24 && blockIdx.x * 8 + (threadIdx.x & 7) < E) {
    const int arr = threadIdx.x >> 3, col = threadIdx.x & 7;
    const float (*src)[9] = arr == 0 ? sg : (arr == 1 ? sb : sc);
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll 8
    for (int k = 0; k < 128; k += 4) { t0 += src[k][col]; t1 += src[k + 1][col]; t2 += src[k + 2][col]; t3 += src[k + 3][col]; }
    const float t = (t0 + t1) + (t2 + t3);
    const int cc = blockIdx.x * 8 + col;
    if (arr == 0) dgamma[cc] = accumulate ? dgamma[cc] + t : t;
    else if (arr == 1) dbeta[cc] = accumulate ? dbeta[cc] + t : t;
    else if (dx_colsum != nullptr) dx_colsum[cc] = accumulate ? dx_colsum[cc] + t : t;
  }
}

// ---------------------------------------------------------------------------------------------
// Column sums of a bf16 matrix (bias gradients of the Linear layers): x [M,F] bf16 -> out[F] (+)= sum_m x[m,:].
// Stage 1: CTA (column block of 64, row chunk r) -> fp32 partial [R][F]; stage 2 folds the R partials in a fixed order.
constexpr int kColsumChunks = 64;

__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int M, int F, int rows_per_chunk, float* __restrict__ partial) {
  __shared__ float2 red[8][33];
  const int cp = threadIdx.x & 31, grp = threadIdx.x >> 5;      // 32 column pairs x 8 row groups
  const int c = blockIdx.x * 64 + cp * 2;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(M, r0 + rows_per_chunk);
  float2 acc = make_float2(0.f, 0.f);
  if (c < F) {
    for (int r = r0 + grp; r < r1; r += 8) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (size_t)r * F + c));
      acc.x += v.x;
      acc.y += v.y;
    }
  }
  red[grp][cp] = acc;
  __syncthreads();
  if (grp == 0 && c < F) {
    float2 t = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { t.x += red[k][cp].x; t.y += red[k][cp].y; }
    *reinterpret_cast<float2*>(partial + (size_t)blockIdx.y * F + c) = t;
  }
}

__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float* __restrict__ partial, int nparts, int F, float* __restrict__ out, int accumulate) {
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

constexpr int kBwdCtas = 296;     // 2 per SM (register limit: ~105 per thread at E = 768)

template <int VEC>
int launch_fwd(const void* x, int x_bf16, const void* res, void* sum_out, const float* gamma, const float* beta, int M, float eps, void* y, int y_bf16, float* mean,
               float* rstd, cudaStream_t st) {
  const unsigned grid = (unsigned)std::min((M + kWarpsPerCta - 1) / kWarpsPerCta, 148 * 8);
  const int T = kWarpsPerCta * 32;
  if (x_bf16) {
    if (y_bf16) layernorm_fwd_kernel<VEC, true, true><<<grid, T, 0, st>>>(x, res, sum_out, gamma, beta, M, eps, y, mean, rstd);
    else layernorm_fwd_kernel<VEC, false, true><<<grid, T, 0, st>>>(x, res, sum_out, gamma, beta, M, eps, y, mean, rstd);
  } else {
    if (y_bf16) layernorm_fwd_kernel<VEC, true, false><<<grid, T, 0, st>>>(x, res, sum_out, gamma, beta, M, eps, y, mean, rstd);
    else layernorm_fwd_kernel<VEC, false, false><<<grid, T, 0, st>>>(x, res, sum_out, gamma, beta, M, eps, y, mean, rstd);
  }
  return acr::check_launch("layernorm_fwd_kernel");
}
template <int VEC>
int launch_bwd(const void* dy, int dy_bf16, const void* dres, const void* x, int x_bf16, const float* mean, const float* rstd, const float* gamma, int M,
               void* dx, float* dgamma, float* dbeta, float* dx_colsum, int accumulate, float* parts, cudaStream_t st) {
  constexpr int E = 128 * VEC;
  const int ctas = M < kBwdCtas * kWarpsPerCta ? (M + kWarpsPerCta - 1) / kWarpsPerCta : kBwdCtas;
  const int rows_per_cta = (M + ctas - 1) / ctas;
  float* pg = parts;
  float* pb = parts + (size_t)kBwdCtas * E;
  float* pc = dx_colsum ? parts + (size_t)2 * kBwdCtas * E : nullptr;
  const int T = kWarpsPerCta * 32;
#define ACR_LN_BWD(DYB, XB, CS)                                                                                              \
  do {                                                                                                                      \
    auto kfn = layernorm_bwd_kernel<VEC, DYB, XB, CS>;                                                                      \
    const size_t smem = (size_t)(CS ? 3 : 2) * kWarpsPerCta * E * sizeof(float);                                            \
    static bool attr_set[64] = {false};       /* per device: the attribute is a per-device property of the kernel */      \
    if (int e_ = acr_attn::set_max_smem(kfn, smem, attr_set)) return e_;                                                    \
    kfn<<<ctas, T, smem, st>>>(dy, dres, x, mean, rstd, gamma, M, rows_per_cta, dx, pg, pb, pc);                            \
  } while (0)
  if (pc != nullptr) {          // column sums of dx: bf16 stream only (the trunk's residual path)
    ACR_REQUIRE(x_bf16 && dy_bf16, ACR_E_INVAL, "acr_layernorm_bwd: dx_colsum needs bf16 x and dy");
    ACR_LN_BWD(true, true, true);
  } else if (x_bf16) {
    if (dy_bf16) ACR_LN_BWD(true, true, false);
    else ACR_LN_BWD(false, true, false);
  } else {
    if (dy_bf16) ACR_LN_BWD(true, false, false);
    else ACR_LN_BWD(false, false, false);
  }
#undef ACR_LN_BWD
  if (int e = acr::check_launch("layernorm_bwd_kernel")) return e;
  layernorm_bwd_finish_kernel<<<(E + 7) / 8, 1024, 0, st>>>(pg, pb, pc, ctas, E, dgamma, dbeta, dx_colsum, accumulate);
  return acr::check_launch("layernorm_bwd_finish_kernel");
}

}  // namespace

extern "C" size_t acr_layernorm_bwd_workspace(int E) { return E > 0 ? (size_t)3 * kBwdCtas * E * sizeof(float) : 0; }

extern "C" int acr_layernorm_fwd(const void* x, int x_is_bf16, const void* residual, void* sum_out, const float* gamma, const float* beta,
                                 int M, int E, float eps, void* y, int y_is_bf16, float* mean, float* rstd, void* stream) {
  ACR_REQUIRE(x && gamma && beta && y && mean && rstd, ACR_E_INVAL, "acr_layernorm_fwd: null pointer");
  ACR_REQUIRE(M > 0 && E > 0 && E % 128 == 0 && E <= 2048, ACR_E_INVAL, "acr_layernorm_fwd: E=%d must be a multiple of 128, <= 2048", E);
  ACR_REQUIRE((residual == nullptr) == (sum_out == nullptr), ACR_E_INVAL, "acr_layernorm_fwd: residual and sum_out go together");
  ACR_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)residual | (uintptr_t)sum_out) & 15) == 0, ACR_E_ALIGN,
              "acr_layernorm_fwd: 16-byte alignment required");
  cudaStream_t st = (cudaStream_t)stream;
  switch (E / 128) {
    case 1: return launch_fwd<1>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 2: return launch_fwd<2>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 3: return launch_fwd<3>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 4: return launch_fwd<4>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 6: return launch_fwd<6>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 8: return launch_fwd<8>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 10: return launch_fwd<10>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 12: return launch_fwd<12>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    case 16: return launch_fwd<16>(x, x_is_bf16, residual, sum_out, gamma, beta, M, eps, y, y_is_bf16, mean, rstd, st);
    default: acr::set_error("acr_layernorm_fwd: E=%d unsupported (128*{1,2,3,4,6,8,10,12,16})", E); return ACR_E_INVAL;
  }
}

extern "C" int acr_layernorm_bwd(const void* dy, int dy_is_bf16, const void* d_residual, const void* x, int x_is_bf16, const float* mean, const float* rstd,
                                 const float* gamma, int M, int E, void* dx, float* dgamma, float* dbeta, float* dx_colsum, int accumulate,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  ACR_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace, ACR_E_INVAL, "acr_layernorm_bwd: null pointer");
  ACR_REQUIRE(M > 0 && E > 0 && E % 128 == 0 && E <= 2048, ACR_E_INVAL, "acr_layernorm_bwd: E=%d must be a multiple of 128, <= 2048", E);
  ACR_REQUIRE(workspace_bytes >= acr_layernorm_bwd_workspace(E), ACR_E_NOMEM, "acr_layernorm_bwd: workspace too small");
  ACR_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma | (uintptr_t)workspace | (uintptr_t)d_residual) & 15) == 0, ACR_E_ALIGN,
              "acr_layernorm_bwd: 16-byte alignment required");
  cudaStream_t st = (cudaStream_t)stream;
  float* parts = (float*)workspace;
  switch (E / 128) {
    case 1: return launch_bwd<1>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 2: return launch_bwd<2>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 3: return launch_bwd<3>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 4: return launch_bwd<4>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 6: return launch_bwd<6>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 8: return launch_bwd<8>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 10: return launch_bwd<10>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 12: return launch_bwd<12>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    case 16: return launch_bwd<16>(dy, dy_is_bf16, d_residual, x, x_is_bf16, mean, rstd, gamma, M, dx, dgamma, dbeta, dx_colsum, accumulate, parts, st);
    default: acr::set_error("acr_layernorm_bwd: E=%d unsupported (128*{1,2,3,4,6,8,10,12,16})", E); return ACR_E_INVAL;
  }
}

extern "C" size_t acr_colsum_workspace(int F) { return F > 0 ? (size_t)kColsumChunks * F * sizeof(float) : 0; }

extern "C" int acr_colsum_bf16(const void* x, int M, int F, float* out, int accumulate, void* workspace, size_t workspace_bytes,
                               void* stream) {
  ACR_REQUIRE(x && out && workspace, ACR_E_INVAL, "acr_colsum_bf16: null pointer");
  ACR_REQUIRE(M > 0 && F > 0 && F % 2 == 0, ACR_E_INVAL, "acr_colsum_bf16: F must be even");
  ACR_REQUIRE(workspace_bytes >= acr_colsum_workspace(F), ACR_E_NOMEM, "acr_colsum_bf16: workspace too small");
  ACR_REQUIRE(((uintptr_t)x & 3) == 0 && ((uintptr_t)workspace & 7) == 0, ACR_E_ALIGN, "acr_colsum_bf16: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = M < kColsumChunks * 8 ? 1 : kColsumChunks;
  const int rows_per_chunk = (M + chunks - 1) / chunks;
  dim3 grid((F + 63) / 64, chunks);
  colsum_bf16_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, M, F, rows_per_chunk, (float*)workspace);
  if (int e = acr::check_launch("colsum_bf16_kernel")) return e;
  colsum_finish_kernel<<<(F + 63) / 64, 256, 0, st>>>((const float*)workspace, chunks, F, out, accumulate);
  return acr::check_launch("colsum_finish_kernel");
}
