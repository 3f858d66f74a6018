// Head of the dense-CRF regulariser (BASELINE configs[3], train_acr_coco.py + myTool.py:825-857 call shape): the probabilities the
// bilateral filter is applied to.  What the step computes from the patch-token logits z [B, P*P, C] (channel-last, as
// cls_head produces them for layer_4[:,1:]):
//     up   = bilinear(z -> S x S, align_corners=False)                       [B, C, S, S]
//     prob = softmax over [0 (background), up_1 .. up_C]                     [B, C+1, S, S]
//     seg  = bilinear(prob -> S/2 x S/2, align_corners=False, scale 0.5)     [B, C+1, S/2, S/2]   (= the 2x2 mean)
// As library calls this is six elementwise passes over 0.5 GB tensors per step at 448x448, C = 80 (forward and backward): 24 of
// the 37.6 ms of the COCO-shaped step.  Here the full-resolution tensors never exist in the forward (one thread per output pixel
// interpolates, normalises and averages its four source pixels from the L1-resident patch logits); the backward recomputes the
// softmax per full-resolution pixel and emits d(up) once, which bilinear_up_bwd_kernel (gather form, no atomics) folds to patch resolution.
// The interpolation follows ATen's upsample_bilinear2d (source index scale*(dst+0.5)-0.5 clamped at 0, lambda order of its kernel).
#include "common.cuh"

namespace {

struct Tap { int i0, i1; float l0, l1; };   // source rows (or columns) and their weights

__device__ __forceinline__ Tap make_tap(int dst, float scale, int n_in) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  Tap t;
  t.i0 = (int)src;
  if (t.i0 > n_in - 1) t.i0 = n_in - 1;
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.l1 = src - (float)t.i0;
  t.l0 = 1.f - t.l1;
  return t;
}

// interpolated logit of class c at a full-resolution pixel whose four source patches start at p00..p11 (rows of C floats)
__device__ __forceinline__ float interp(const float* p00, const float* p01, const float* p10, const float* p11, const Tap& ty, const Tap& tx, int c) {
  return ty.l0 * (tx.l0 * __ldg(p00 + c) + tx.l1 * __ldg(p01 + c)) + ty.l1 * (tx.l0 * __ldg(p10 + c) + tx.l1 * __ldg(p11 + c));
}

struct Pix {      // one full-resolution pixel: its four patch rows, taps and softmax normaliser
  const float *p00, *p01, *p10, *p11;
  Tap ty, tx;
  float m, inv_s;
};

__device__ __forceinline__ void pix_setup(Pix& q, const float* zb, int Y, int X, int P, int C, float scale) {
  q.ty = make_tap(Y, scale, P);
  q.tx = make_tap(X, scale, P);
  q.p00 = zb + ((long long)q.ty.i0 * P + q.tx.i0) * C;
  q.p01 = zb + ((long long)q.ty.i0 * P + q.tx.i1) * C;
  q.p10 = zb + ((long long)q.ty.i1 * P + q.tx.i0) * C;
  q.p11 = zb + ((long long)q.ty.i1 * P + q.tx.i1) * C;
  // online max / sum over [0, up_1..up_C]
  float m = 0.f, s = 1.f;                 // the background logit is the constant 0
  for (int c = 0; c < C; ++c) {
    const float v = interp(q.p00, q.p01, q.p10, q.p11, q.ty, q.tx, c);
    if (v > m) { s = s * __expf(m - v) + 1.f; m = v; } else s += __expf(v - m);
  }
  q.m = m;
  q.inv_s = 1.f / s;
}

// seg[b,k,y,x] = mean over the 2x2 full-resolution pixels of softmax_k.  One thread per (b,y,x).
__global__ void __launch_bounds__(128)
crf_head_fwd_kernel(const float* __restrict__ z, int P, int C, int S, float scale, float* __restrict__ seg) {
  const int h = S / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= h) return;
  const float* zb = z + (long long)b * P * P * C;
  Pix q[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) pix_setup(q[u], zb, 2 * y + (u >> 1), 2 * x + (u & 1), P, C, scale);
  float* out = seg + ((long long)b * (C + 1) * h + y) * h + x;
  const long long plane = (long long)h * h;
  float bg = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) bg += __expf(-q[u].m) * q[u].inv_s;
  out[0] = 0.25f * bg;
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += __expf(interp(q[u].p00, q[u].p01, q[u].p10, q[u].p11, q[u].ty, q[u].tx, c) - q[u].m) * q[u].inv_s;
    out[(long long)(c + 1) * plane] = 0.25f * acc;
  }
}

// d_up[b,c,Y,X] = p_c * (g_c - sum_j p_j g_j) with p = softmax at the full-resolution pixel and g = 0.25 * g_seg[b,:,Y/2,X/2]
// (the 2x2 mean gives every source pixel a quarter of the output gradient).  One thread per (b,Y,X).
__global__ void __launch_bounds__(128)
crf_head_bwd_kernel(const float* __restrict__ z, const float* __restrict__ g_seg, int P, int C, int S, float scale, float* __restrict__ d_up) {
  const int h = S / 2;
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, b = blockIdx.z;
  if (X >= S) return;
  Pix q;
  pix_setup(q, z + (long long)b * P * P * C, Y, X, P, C, scale);
  const long long plane = (long long)h * h;
  const float* g = g_seg + ((long long)b * (C + 1) * h + (Y >> 1)) * h + (X >> 1);
  float t = __expf(-q.m) * q.inv_s * __ldg(g);      // background term of sum_j p_j g_j
  for (int c = 0; c < C; ++c)
    t += __expf(interp(q.p00, q.p01, q.p10, q.p11, q.ty, q.tx, c) - q.m) * q.inv_s * __ldg(g + (long long)(c + 1) * plane);
  float* out = d_up + ((long long)b * C * S + Y) * S + X;
  const long long oplane = (long long)S * S;
  for (int c = 0; c < C; ++c) {
    const float p = __expf(interp(q.p00, q.p01, q.p10, q.p11, q.ty, q.tx, c) - q.m) * q.inv_s;
    out[(long long)c * oplane] = 0.25f * p * (__ldg(g + (long long)(c + 1) * plane) - t);
  }
}

// Backward of the bilinear up-sampling (align_corners=False): d_patch[b,c,py,px] = sum_{Y,X} wy(Y,py) wx(X,px) d_up[b,c,Y,X].
// (ATen's upsample_bilinear2d_backward scatters with one atomic per output pixel and tap: 16.6 ms for [8,80,448,448].)
// One CTA per (b,c) plane, one thread per column X walking the rows once.  Rows with the same upper source row i0 form a group;
// when a group ends, patch row py = i0 has all its contributions (weight l0 from this group, l1 from the previous one), the
// column sums go to shared memory and 16 lanes per patch column fold them with the column weights.
constexpr int kUpBwdThreads = 512;
__global__ void __launch_bounds__(kUpBwdThreads)
bilinear_up_bwd_kernel(const float* __restrict__ d_up, int P, int S, float scale, float* __restrict__ d_patch) {
  extern __shared__ float colsum[];          // [S]
  const float* plane = d_up + (size_t)blockIdx.x * S * S;
  float* outp = d_patch + (size_t)blockIdx.x * P * P;
  const int X = threadIdx.x;                 // (S <= kUpBwdThreads)
  float carry = 0.f;
  int Y = 0;
  while (Y < S) {
    const int g = make_tap(Y, scale, P).i0;
    int Yend = Y + 1;
    while (Yend < S && make_tap(Yend, scale, P).i0 == g) ++Yend;
    float a0 = carry, a1 = 0.f;
    if (X < S) {
#pragma unroll 4
      for (int yy = Y; yy < Yend; ++yy) {
        const Tap t = make_tap(yy, scale, P);
        const float v = __ldg(plane + (size_t)yy * S + X);
        a0 = fmaf(t.l0, v, a0);
        if (t.i1 != t.i0) a1 = fmaf(t.l1, v, a1); else a0 = fmaf(t.l1, v, a0);
      }
      colsum[X] = a0;
    }
    carry = a1;
    Y = Yend;
    __syncthreads();
    // patch row g: 16 lanes per patch column
    for (int pxb = (threadIdx.x >> 5) * 2; pxb < P; pxb += (kUpBwdThreads >> 5) * 2) {      // warp-uniform trip count (shuffles below)
      const int px = pxb + ((threadIdx.x >> 4) & 1);
      float acc = 0.f;
      for (int xx = threadIdx.x & 15; xx < S; xx += 16) {
        const Tap t = make_tap(xx, scale, P);
        const float w = (t.i0 == px ? t.l0 : 0.f) + (t.i1 == px ? t.l1 : 0.f);
        acc = fmaf(w, colsum[xx], acc);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, 16);
      if ((threadIdx.x & 15) == 0 && px < P) outp[g * P + px] = acc;
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int acr_bilinear_up_bwd(const float* d_up, int planes, int P, int S, float* d_patch, void* stream) {
  ACR_REQUIRE(d_up && d_patch, ACR_E_INVAL, "acr_bilinear_up_bwd: null pointer");
  ACR_REQUIRE(planes > 0 && P > 0 && S >= P && S <= kUpBwdThreads, ACR_E_INVAL, "acr_bilinear_up_bwd: bad shape (P <= S <= %d)", kUpBwdThreads);
  bilinear_up_bwd_kernel<<<planes, kUpBwdThreads, (size_t)S * sizeof(float), (cudaStream_t)stream>>>(d_up, P, S, (float)P / (float)S, d_patch);
  return acr::check_launch("bilinear_up_bwd_kernel");
}

extern "C" int acr_crf_head_fwd(const float* logits, int B, int P, int C, int S, float* seg, void* stream) {
  ACR_REQUIRE(logits && seg, ACR_E_INVAL, "acr_crf_head_fwd: null pointer");
  ACR_REQUIRE(B > 0 && B <= 65535 && P > 0 && C > 0 && S >= 2 && (S % 2) == 0 && S / 2 <= 65535, ACR_E_INVAL, "acr_crf_head_fwd: bad shape (S even)");
  const int h = S / 2;
  crf_head_fwd_kernel<<<dim3((h + 127) / 128, h, B), 128, 0, (cudaStream_t)stream>>>(logits, P, C, S, (float)P / (float)S, seg);
  return acr::check_launch("crf_head_fwd_kernel");
}

extern "C" int acr_crf_head_bwd(const float* logits, const float* g_seg, int B, int P, int C, int S, float* d_up, void* stream) {
  ACR_REQUIRE(logits && g_seg && d_up, ACR_E_INVAL, "acr_crf_head_bwd: null pointer");
  ACR_REQUIRE(B > 0 && B <= 65535 && P > 0 && C > 0 && S >= 2 && (S % 2) == 0 && S <= 65535, ACR_E_INVAL, "acr_crf_head_bwd: bad shape (S even)");
  crf_head_bwd_kernel<<<dim3((S + 127) / 128, S, B), 128, 0, (cudaStream_t)stream>>>(logits, g_seg, P, C, S, (float)P / (float)S, d_up);
  return acr::check_launch("crf_head_bwd_kernel");
}
