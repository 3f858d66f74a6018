// (a7) All-pairs consistency loss: forward + gradient in ONE streaming pass over the two
// [B,L,N,N] head-mean stacks.  Reference: train_acr.py:143-161 (slices, 3*p in-place flips,
// two F.l1_loss).  The flips are the index permutation pi(1+r*p+c) = 1+r*p+(p-1-c), pi(0)=0,
// applied to rows and columns of view 2; it is done in index math here, so
// neither input is modified and nothing is copied.
//
// HBM-bound: compulsory traffic = read A1, A2 once + write G1, G2 once = 16*B*L*N*N bytes.
// One warp per (b,l,i) row: row i of A1 is paired with row pi(i) of A2; column j with pi(j).
// Per-row partial |d| sums go to a scratch array and are folded by a second, single-CTA kernel in a
// fixed order (deterministic; no float atomics).
#include "common.cuh"
#include <type_traits>

namespace {


// read-only load as a volatile asm statement: the compiler keeps these in program order (see the barrier in the row loop)
__device__ __forceinline__ float ld_nc(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ int flip_token(int t, int p) {
  if (t == 0) return 0;
  const int q = t - 1;
  const int r = q / p;
  const int c = q - r * p;
  return 1 + r * p + (p - 1 - c);
}

// kGrad: 0 = loss only, 1 = dense fp32 gradients, 2 = sign codes (one byte per element: 0x00 zero, 0x3F plus, 0xBF minus,
// i.e. the top byte of +-0.5f, so that a consumer decodes with one byte-permute: float(code << 24) * 2w).
// One WARP per (b,l,i) row at a time, kWarps warps per CTA walking kRowsPerWarp rows each; each lane keeps kUnroll
// independent load pairs in flight.  The column permutation pi is the same for every row: it is tabulated once per CTA in
// shared memory (the two integer divisions of flip_token per element made the kernel issue bound: 76 % of the issue slots
// at 54 % of HBM peak, profiles/r01g_ncu_refine_kernels.txt).
constexpr int kWarps = 8;
constexpr int kRowsPerWarp = 4;
constexpr int kRowsPerCta = kWarps * kRowsPerWarp;
constexpr int kUnroll = 8;

template <int kGrad>
__global__ void __launch_bounds__(kWarps * 32)
consistency_rows_kernel(const float* __restrict__ a1, const float* __restrict__ a2, long long rows,
                        int N, int p, float w_cls, float w_aff,
                        float* __restrict__ g1, float* __restrict__ g2, long long g_ld,
                        unsigned char* __restrict__ c1, unsigned char* __restrict__ c2, long long c_ld,
                        float* __restrict__ partials) {
  extern __shared__ unsigned short pi_tab[];      // pi(j), j < N
  __shared__ float s_part[kWarps][2];             // per warp: sum |d| of its cls rows (i = 0) and of its patch rows
  for (int j = threadIdx.x; j < N; j += blockDim.x) pi_tab[j] = (unsigned short)flip_token(j, p);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float sum_cls = 0.f, sum_aff = 0.f;
  for (int rw = 0; rw < kRowsPerWarp; ++rw) {
  const long long row = (long long)blockIdx.x * kRowsPerCta + (long long)rw * kWarps + (threadIdx.x >> 5);      // (b*L + l)*N + i
  if (row >= rows) break;
  const int i = (int)(row % N);
  const long long img = row / N;               // b*L + l
  const int pi_i = pi_tab[i];
  const float* r1 = a1 + (img * N + i) * (long long)N;
  const float* r2 = a2 + (img * N + pi_i) * (long long)N;
  const float w = (i == 0) ? w_cls : w_aff;
  float* o1 = (kGrad == 1) ? g1 + (img * N + i) * g_ld : nullptr;
  float* o2 = (kGrad == 1) ? g2 + (img * N + pi_i) * g_ld : nullptr;
  unsigned char* b1 = (kGrad == 2) ? c1 + (img * N + i) * c_ld : nullptr;
  unsigned char* b2 = (kGrad == 2) ? c2 + (img * N + pi_i) * c_ld : nullptr;

  float acc = 0.f;
  // full chunks of 32 * kUnroll columns run without bounds checks; the tail chunk (N = p*p+1) with them
  auto chunk = [&](int j0, auto full_tag) {
    constexpr bool kFull = decltype(full_tag)::value;
    float v1[kUnroll], v2[kUnroll];
    int pj[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int j = j0 + k * 32 + lane;
      if (kFull || j < N) {
        pj[k] = pi_tab[j];
        v1[k] = ld_nc(r1 + j);
        v2[k] = ld_nc(r2 + pj[k]);
      }
    }
    // keep all 2 * kUnroll loads in flight before the first store: ptxas otherwise sinks every load next to its use
    // (load, compare, store, next load ...: 2.1 instead of 3.3 TB/s); a warp barrier is a scheduling fence for memory operations
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int j = j0 + k * 32 + lane;
      if (kFull || j < N) {
        float s = 0.f;
        unsigned char code = 0, ncode = 0;
        if (j > 0) {
          const float d = v1[k] - v2[k];
          acc += fabsf(d);
          s = (d > 0.f) ? w : ((d < 0.f) ? -w : 0.f);
          code = (d > 0.f) ? 0x3F : ((d < 0.f) ? 0xBF : 0x00);
          ncode = (d > 0.f) ? 0xBF : ((d < 0.f) ? 0x3F : 0x00);
        }
        if (kGrad == 1) {
          o1[j] = s;
          o2[pj[k]] = -s;
        }
        if (kGrad == 2) {
          b1[j] = code;
          b2[pj[k]] = ncode;
        }
      }
    }
  };
  int j0 = 0;
  for (; j0 + 32 * kUnroll <= N; j0 += 32 * kUnroll) chunk(j0, std::true_type{});
  if (j0 < N) chunk(j0, std::false_type{});
  if (kGrad == 2) {      // row padding [N, c_ld): consumers read whole 16-byte chunks of a row; keep them defined (= no gradient)
    for (long long j = N + lane; j < c_ld; j += 32) {
      b1[j] = 0;
      b2[j] = 0;
    }
  }
  acc = acr::warp_sum(acc);
  if (i == 0) sum_cls += acc; else sum_aff += acc;
  }
  // one (cls, aff) pair per CTA, summed in a fixed order: the finish kernel folds 2 * gridDim.x numbers instead of one per row
  if (lane == 0) { s_part[threadIdx.x >> 5][0] = sum_cls; s_part[threadIdx.x >> 5][1] = sum_aff; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) v += s_part[w][threadIdx.x];
    partials[2 * (long long)blockIdx.x + threadIdx.x] = v;
  }
}

// Folds the per-CTA (cls, aff) partials in a fixed order (deterministic; no float atomics).
__global__ void __launch_bounds__(1024)
consistency_finish_kernel(const float2* __restrict__ partials, int nparts,
                          double inv_cls, double inv_aff, float* __restrict__ loss2) {
  __shared__ double s_cls[32], s_aff[32];
  double c = 0.0, a = 0.0;
  for (int r = threadIdx.x; r < nparts; r += blockDim.x) {
    const float2 v = partials[r];
    c += (double)v.x;
    a += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    a += __shfl_xor_sync(0xffffffffu, a, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s_cls[w] = c; s_aff[w] = a; }
  __syncthreads();
  if (w == 0) {
    c = (lane < (int)(blockDim.x >> 5)) ? s_cls[lane] : 0.0;
    a = (lane < (int)(blockDim.x >> 5)) ? s_aff[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, o);
      a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    if (lane == 0) {
      loss2[0] = (float)(c * inv_cls);
      loss2[1] = (float)(a * inv_aff);
    }
  }
}

}  // namespace

extern "C" size_t acr_consistency_workspace(int B, int L, int N) {
  if (B <= 0 || L <= 0 || N <= 0) return 0;
  return acr::align_up((size_t)B * L * N * sizeof(float), 256);
}

extern "C" int acr_consistency_fwd_bwd(const float* attn1, const float* attn2, int B, int L, int N, int p,
                                       float alpha_cls, float alpha_aff,
                                       float* loss2, float* g1, float* g2, long long g_row_stride,
                                       unsigned char* code1, unsigned char* code2, long long code_row_stride,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  ACR_REQUIRE(attn1 && attn2 && loss2 && workspace, ACR_E_INVAL, "acr_consistency_fwd_bwd: null pointer");
  ACR_REQUIRE(B > 0 && L > 0 && p > 0, ACR_E_INVAL, "acr_consistency_fwd_bwd: bad B/L/p");
  ACR_REQUIRE(N == p * p + 1, ACR_E_INVAL, "acr_consistency_fwd_bwd: N=%d is not p*p+1 (p=%d)", N, p);
  ACR_REQUIRE((g1 == nullptr) == (g2 == nullptr), ACR_E_INVAL, "acr_consistency_fwd_bwd: g1/g2 must both be set or null");
  ACR_REQUIRE(g1 == nullptr || g_row_stride >= N, ACR_E_INVAL, "acr_consistency_fwd_bwd: g_row_stride < N");
  ACR_REQUIRE((code1 == nullptr) == (code2 == nullptr), ACR_E_INVAL, "acr_consistency_fwd_bwd: code1/code2 must both be set or null");
  ACR_REQUIRE(code1 == nullptr || g1 == nullptr, ACR_E_INVAL, "acr_consistency_fwd_bwd: ask for dense gradients OR sign codes");
  ACR_REQUIRE(code1 == nullptr || code_row_stride >= N, ACR_E_INVAL, "acr_consistency_fwd_bwd: code_row_stride < N");
  ACR_REQUIRE(workspace_bytes >= acr_consistency_workspace(B, L, N), ACR_E_NOMEM,
              "acr_consistency_fwd_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * L * N;
  ACR_REQUIRE(rows < (1ll << 31), ACR_E_INVAL, "acr_consistency_fwd_bwd: too many rows");
  ACR_REQUIRE(N <= 16384, ACR_E_INVAL, "acr_consistency_fwd_bwd: N=%d too large (<= 16384)", N);
  const size_t smem = (size_t)N * sizeof(unsigned short);
  const double cnt_cls = (double)B * L * (N - 1);
  const double cnt_aff = (double)B * L * (double)(N - 1) * (double)(N - 1);
  const float w_cls = (float)((double)alpha_cls / cnt_cls);
  const float w_aff = (float)((double)alpha_aff / cnt_aff);
  float* partials = (float*)workspace;
  const unsigned grid = (unsigned)((rows + kRowsPerCta - 1) / kRowsPerCta);
  if (g1) {
    consistency_rows_kernel<1><<<grid, kWarps * 32, smem, st>>>(attn1, attn2, rows, N, p, w_cls, w_aff, g1, g2, g_row_stride, nullptr, nullptr, 0, partials);
  } else if (code1) {
    consistency_rows_kernel<2><<<grid, kWarps * 32, smem, st>>>(attn1, attn2, rows, N, p, w_cls, w_aff, nullptr, nullptr, 0, code1, code2, code_row_stride, partials);
  } else {
    consistency_rows_kernel<0><<<grid, kWarps * 32, smem, st>>>(attn1, attn2, rows, N, p, w_cls, w_aff, nullptr, nullptr, 0, nullptr, nullptr, 0, partials);
  }
  if (int e = acr::check_launch("consistency_rows_kernel")) return e;
  consistency_finish_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const float2*>(partials), (int)grid, 1.0 / cnt_cls, 1.0 / cnt_aff, loss2);
  return acr::check_launch("consistency_finish_kernel");
}
