// (a11) Permutohedral-lattice bilateral filter on the GPU.
// Reference: wrapper/bilateralfilter/bilateralfilter.cpp:4-55 (feature build, per-image / per-plane loop)
// and permutohedral.cpp:115-440 (Permutohedral::init, SSE branch) / :507-631 (Permutohedral::compute).
//
// What must match the reference (it is a lattice APPROXIMATION of a 5-D Gaussian, not an exact one):
//   * features (x/sxy, y/sxy, r/srgb, g/srgb, b/srgb), d = 5 (bilateralfilter.cpp:8-16);
//   * elevation with scale_i = (d+1)sqrt(2/3)/sqrt((i+1)(i+2)) by the running-sum recurrence (:339-345),
//     separate multiply and add roundings (the SSE code has no FMA -> __fmul_rn/__fadd_rn here);
//   * nearest-EVEN rounding to the remainder-0 simplex (:349-357, _mm_round_ps / cvtps) -> rintf;
//   * rank by pairwise comparison, sum fix-up (:359-380), barycentric weights in the same accumulation
//     order (:383-403), the d+1 vertex keys rem0 + canonical[remainder][rank] (:404-413);
//   * blur: for axis j = 0..d sequentially new = old + 0.5 (old[n1] + old[n2]) with n1/n2 = key -/+ 1 on all
//     coordinates and +/- d on coordinate j, missing neighbour = 0 (:423-436, :548-566);
//   * slice with alpha = 1/(1+2^-d), no normalisation (:567-583).
// What is free: vertex numbering and hash function (results are indexed by key, not by id), the order of
// the splat additions (float atomics here; differences are ~1e-7 relative), and filtering all K planes in one
// pass (the planes are independent and the filter is linear; the reference loops K times with value_size=1).
//
// Data layout in HBM (per image, carved from the caller's workspace):
//   ckey  [6*HW] 3 x u32   packed int16 key of every (pixel, remainder) candidate
//   bary  [6*HW] f32       barycentric weight of the candidate
//   table [cap]  i32       open-addressing hash: representative candidate index or -1 (cap = pow2 >= 12*HW)
//   rep   [6*HW] i32       candidate -> representative candidate; later overwritten by vertex id + 1
//   vid   [6*HW] i32       representative candidate -> dense vertex id
//   vcand [6*HW] i32       vertex id -> representative candidate
//   nbr   [6][6*HW] int2   blur neighbours (vertex id + 1; 0 = missing)
//   val0/val1 [(6*HW+1) * K] f32   lattice values, vertex-major with the K planes contiguous; slot 0 = zeros
#include "common.cuh"
#include <cmath>
#include <vector>
#include <type_traits>

namespace {

constexpr int D5 = 5;          // feature dimension
constexpr int D6 = D5 + 1;

struct Key { uint32_t a, b, c; };   // k0|k1<<16, k2|k3<<16, k4

template <typename T>
__device__ __forceinline__ T* img_ptr(T* base, size_t stride_bytes, int n) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(const_cast<typename std::remove_const<T>::type*>(base)) + stride_bytes * n);
}

struct Scales { float s[D5]; };

__device__ __forceinline__ Key pack_key(const short* k) {
  Key r;
  r.a = (uint32_t)(uint16_t)k[0] | ((uint32_t)(uint16_t)k[1] << 16);
  r.b = (uint32_t)(uint16_t)k[2] | ((uint32_t)(uint16_t)k[3] << 16);
  r.c = (uint32_t)(uint16_t)k[4];
  return r;
}
__device__ __forceinline__ void unpack_key(const Key& r, short* k) {
  k[0] = (short)(r.a & 0xffff); k[1] = (short)(r.a >> 16);
  k[2] = (short)(r.b & 0xffff); k[3] = (short)(r.b >> 16);
  k[4] = (short)(r.c & 0xffff);
}
__device__ __forceinline__ bool key_eq(const Key& x, const Key& y) { return x.a == y.a && x.b == y.b && x.c == y.c; }
__device__ __forceinline__ uint32_t key_hash(const short* k) {
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < D5; ++i) { r += (uint64_t)(int64_t)k[i]; r *= 1664525ull; }
  return (uint32_t)(r ^ (r >> 32));
}

// Step 1: per pixel, lattice coordinates -> 6 candidate keys + barycentric weights.
__global__ void __launch_bounds__(256)
lattice_embed_kernel(const float* __restrict__ image, int H, int W, int HWpad, float sigmargb, float sigmaxy, Scales sc,
                     Key* __restrict__ ckey, float* __restrict__ bary_out, size_t ws) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (idx >= HWpad) return;
  image += (size_t)blockIdx.z * 3 * HW;
  ckey = img_ptr(ckey, ws, blockIdx.z);
  bary_out = img_ptr(bary_out, ws, blockIdx.z);
  const int px = idx % W, py = idx / W;
  float f[D5];
  if (idx < HW) {
    f[0] = __fdiv_rn((float)px, sigmaxy);
    f[1] = __fdiv_rn((float)py, sigmaxy);
    f[2] = __fdiv_rn(__ldg(image + idx), sigmargb);
    f[3] = __fdiv_rn(__ldg(image + HW + idx), sigmargb);
    f[4] = __fdiv_rn(__ldg(image + 2 * HW + idx), sigmargb);
  } else {
    // The reference embeds pixels four at a time and pads the last group with all-zero feature vectors
    // (permutohedral.cpp:182-186); those phantom pixels still create lattice vertices (:404-413), which
    // take part in the blur.  Reproduce them: keys are inserted, nothing is splatted or sliced.
#pragma unroll
    for (int i = 0; i < D5; ++i) f[i] = 0.f;
  }

  float elevated[D6];
  float sm = 0.f;
#pragma unroll
  for (int j = D5; j > 0; --j) {
    const float cf = __fmul_rn(f[j - 1], sc.s[j - 1]);
    elevated[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
    sm = __fadd_rn(sm, cf);
  }
  elevated[0] = sm;

  const float invdplus1 = 1.0f / (float)D6;
  const float dplus1 = (float)D6;
  float rem0[D6], rank[D6];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < D6; ++i) {
    const float v = rintf(__fmul_rn(invdplus1, elevated[i]));
    rem0[i] = __fmul_rn(v, dplus1);
    sum = __fadd_rn(sum, v);
    rank[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < D5; ++i) {
    const float di = __fsub_rn(elevated[i], rem0[i]);
#pragma unroll
    for (int j = i + 1; j < D6; ++j) {
      const float dj = __fsub_rn(elevated[j], rem0[j]);
      const float c = (di < dj) ? 1.f : 0.f;
      rank[i] += c;
      rank[j] += 1.f - c;
    }
  }
#pragma unroll
  for (int i = 0; i < D6; ++i) {
    rank[i] += sum;
    const float add = (rank[i] < 0.f) ? dplus1 : 0.f;
    const float sub = (rank[i] >= dplus1) ? dplus1 : 0.f;
    rank[i] += add - sub;
    rem0[i] += add - sub;
  }
  float bary[D6 + 1];
#pragma unroll
  for (int i = 0; i < D6 + 1; ++i) bary[i] = 0.f;
#pragma unroll
  for (int i = 0; i < D6; ++i) {
    const float v = __fmul_rn(__fsub_rn(elevated[i], rem0[i]), invdplus1);
    const int p = D5 - (int)rank[i];
    // dynamic index into a register array: unrolled select keeps it in registers
#pragma unroll
    for (int q = 0; q < D6 + 1; ++q) {
      if (q == p) bary[q] = __fadd_rn(bary[q], v);
      if (q == p + 1) bary[q] = __fsub_rn(bary[q], v);
    }
  }
  bary[0] = __fadd_rn(bary[0], __fadd_rn(1.0f, bary[D6]));

#pragma unroll
  for (int r = 0; r < D6; ++r) {
    short key[D5];
#pragma unroll
    for (int i = 0; i < D5; ++i) {
      const int rk = (int)rank[i];
      const int canon = (rk <= D5 - r) ? r : r - D6;     // canonical[r*(d+1) + rk]
      key[i] = (short)((int)rem0[i] + canon);
    }
    ckey[(size_t)idx * D6 + r] = pack_key(key);
    bary_out[(size_t)idx * D6 + r] = bary[r];
  }
}

// Step 2: insert every candidate; the CAS winner of a slot becomes the representative and draws a dense id.
__global__ void __launch_bounds__(256)
lattice_insert_kernel(const Key* __restrict__ ckey, int ncand, int* __restrict__ table, uint32_t mask,
                      int* __restrict__ rep, int* __restrict__ vid, int* __restrict__ vcand, int* __restrict__ counter, size_t ws) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  ckey = img_ptr(ckey, ws, blockIdx.z); table = img_ptr(table, ws, blockIdx.z); rep = img_ptr(rep, ws, blockIdx.z);
  vid = img_ptr(vid, ws, blockIdx.z); vcand = img_ptr(vcand, ws, blockIdx.z); counter = img_ptr(counter, ws, blockIdx.z);
  const Key k = ckey[c];
  short ks[D5];
  unpack_key(k, ks);
  uint32_t h = key_hash(ks) & mask;
  while (true) {
    int e = ((volatile int*)table)[h];
    if (e == -1) {
      e = atomicCAS(&table[h], -1, c);
      if (e == -1) {
        const int id = atomicAdd(counter, 1);
        vid[c] = id;
        vcand[id] = c;
        rep[c] = c;
        return;
      }
    }
    if (key_eq(ckey[e], k)) { rep[c] = e; return; }
    h = (h + 1) & mask;
  }
}

// Step 3: candidate -> (vertex id + 1).
__global__ void __launch_bounds__(256)
lattice_offset_kernel(int* __restrict__ rep, const int* __restrict__ vid, int ncand, size_t ws) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  rep = img_ptr(rep, ws, blockIdx.z); vid = img_ptr(vid, ws, blockIdx.z);
  rep[c] = vid[rep[c]] + 1;
}

__device__ __forceinline__ int lattice_find(const Key* __restrict__ ckey, const int* __restrict__ table, uint32_t mask,
                                            const int* __restrict__ vid, const short* ks) {
  const Key k = pack_key(ks);
  uint32_t h = key_hash(ks) & mask;
  while (true) {
    const int e = table[h];
    if (e == -1) return 0;
    if (key_eq(ckey[e], k)) return vid[e] + 1;
    h = (h + 1) & mask;
  }
}

// Step 4: blur neighbours of every vertex along each of the d+1 axes.  One thread per (vertex, axis).
__global__ void __launch_bounds__(256)
lattice_neighbors_kernel(const Key* __restrict__ ckey, const int* __restrict__ table, uint32_t mask,
                         const int* __restrict__ vid, const int* __restrict__ vcand, const int* __restrict__ counter,
                         int stride, int2* __restrict__ nbr, size_t ws) {
  ckey = img_ptr(ckey, ws, blockIdx.z); table = img_ptr(table, ws, blockIdx.z); vid = img_ptr(vid, ws, blockIdx.z);
  vcand = img_ptr(vcand, ws, blockIdx.z); counter = img_ptr(counter, ws, blockIdx.z); nbr = img_ptr(nbr, ws, blockIdx.z);
  const int M = *counter;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = t / D6, j = t % D6;
  if (v >= M) return;
  short key[D5], n1[D5], n2[D5];
  unpack_key(ckey[vcand[v]], key);
#pragma unroll
  for (int k = 0; k < D5; ++k) { n1[k] = key[k] - 1; n2[k] = key[k] + 1; }
#pragma unroll
  for (int k = 0; k < D5; ++k)
    if (k == j) { n1[k] = key[k] + D5; n2[k] = key[k] - D5; }
  int2 r;
  r.x = lattice_find(ckey, table, mask, vid, n1);
  r.y = lattice_find(ckey, table, mask, vid, n2);
  nbr[(size_t)j * stride + v] = r;
}

__global__ void __launch_bounds__(256)
zero_values_kernel(float* __restrict__ v0, float* __restrict__ v1, const int* __restrict__ counter, int K, size_t ws) {
  v0 = img_ptr(v0, ws, blockIdx.z); v1 = img_ptr(v1, ws, blockIdx.z); counter = img_ptr(counter, ws, blockIdx.z);
  const long long n = ((long long)(*counter) + 1) * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    v0[i] = 0.f;
    v1[i] = 0.f;
  }
}

constexpr int kPixTile = 32;

// Splat: block = 32 pixels; planes staged through smem so that both the plane-major input reads and the
// vertex-major value updates are coalesced.
__global__ void __launch_bounds__(256)
lattice_splat_kernel(const float* __restrict__ in, int K, int HW, const int* __restrict__ offs,
                     const float* __restrict__ bary, float* __restrict__ values, size_t ws) {
  extern __shared__ float tile[];   // [kPixTile][K+1]
  in += (size_t)blockIdx.z * K * HW;
  offs = img_ptr(offs, ws, blockIdx.z); bary = img_ptr(bary, ws, blockIdx.z); values = img_ptr(values, ws, blockIdx.z);
  const int p0 = blockIdx.x * kPixTile;
  const int KP = K + 1;
  for (int e = threadIdx.x; e < K * kPixTile; e += blockDim.x) {
    const int k = e / kPixTile, p = e % kPixTile;
    tile[p * KP + k] = (p0 + p < HW) ? __ldg(in + (size_t)k * HW + p0 + p) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = w; p < kPixTile; p += nw) {
    const int pix = p0 + p;
    if (pix >= HW) break;
#pragma unroll
    for (int r = 0; r < D6; ++r) {
      const int o = __ldg(offs + (size_t)pix * D6 + r);
      const float wt = __ldg(bary + (size_t)pix * D6 + r);
      float* dst = values + (size_t)o * K;
      for (int k = lane; k < K; k += 32) atomicAdd(dst + k, __fmul_rn(wt, tile[p * KP + k]));
    }
  }
}

// One blur pass along axis j: new = old + 0.5 * (old[n1] + old[n2]).  One thread per (vertex, plane).
__global__ void __launch_bounds__(256)
lattice_blur_kernel(const float* __restrict__ oldv, float* __restrict__ newv, const int2* __restrict__ nbr,
                    const int* __restrict__ counter, int K, size_t ws) {
  oldv = img_ptr(oldv, ws, blockIdx.z); newv = img_ptr(newv, ws, blockIdx.z); nbr = img_ptr(nbr, ws, blockIdx.z);
  counter = img_ptr(counter, ws, blockIdx.z);
  const long long M = *counter;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < M * K; t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t / K), k = (int)(t % K);
    const int2 nb = __ldg(nbr + v);
    const float a = oldv[(size_t)nb.x * K + k], b = oldv[(size_t)nb.y * K + k];
    newv[(size_t)(v + 1) * K + k] = __fadd_rn(oldv[(size_t)(v + 1) * K + k], __fmul_rn(0.5f, __fadd_rn(a, b)));
  }
}

__global__ void __launch_bounds__(256)
lattice_slice_kernel(const float* __restrict__ values, int K, int HW, const int* __restrict__ offs,
                     const float* __restrict__ bary, float alpha, float* __restrict__ out, size_t ws) {
  extern __shared__ float tile[];   // [kPixTile][K+1]
  out += (size_t)blockIdx.z * K * HW;
  values = img_ptr(values, ws, blockIdx.z); offs = img_ptr(offs, ws, blockIdx.z); bary = img_ptr(bary, ws, blockIdx.z);
  const int p0 = blockIdx.x * kPixTile;
  const int KP = K + 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = w; p < kPixTile; p += nw) {
    const int pix = p0 + p;
    if (pix >= HW) break;
    int o[D6];
    float wt[D6];
#pragma unroll
    for (int r = 0; r < D6; ++r) {
      o[r] = __ldg(offs + (size_t)pix * D6 + r);
      wt[r] = __fmul_rn(__ldg(bary + (size_t)pix * D6 + r), alpha);
    }
    for (int k = lane; k < K; k += 32) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < D6; ++r) s = __fadd_rn(s, __fmul_rn(wt[r], values[(size_t)o[r] * K + k]));
      tile[p * KP + k] = s;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * kPixTile; e += blockDim.x) {
    const int k = e / kPixTile, p = e % kPixTile;
    if (p0 + p < HW) out[(size_t)k * HW + p0 + p] = tile[p * KP + k];
  }
}

struct Carve {
  Key* ckey; float* bary; int* table; int* rep; int* vid; int* vcand; int2* nbr; float* val0; float* val1; int* counter;
  uint32_t cap;
  size_t total;
};

Carve carve(void* base, int K, int H, int W) {
  Carve c{};
  const size_t HW = ((size_t)H * W + 3) / 4 * 4, nc = HW * D6;   // incl. the reference's phantom pad pixels
  uint32_t cap = 1;
  while (cap < 2 * nc) cap <<= 1;
  c.cap = cap;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += acr::align_up(bytes, 256); return (char*)base + o; };
  c.counter = (int*)take(256);
  c.ckey = (Key*)take(nc * sizeof(Key));
  c.bary = (float*)take(nc * sizeof(float));
  c.table = (int*)take((size_t)cap * sizeof(int));
  c.rep = (int*)take(nc * sizeof(int));
  c.vid = (int*)take(nc * sizeof(int));
  c.vcand = (int*)take(nc * sizeof(int));
  c.nbr = (int2*)take(nc * D6 * sizeof(int2));
  c.val0 = (float*)take((nc + 1) * K * sizeof(float));
  c.val1 = (float*)take((nc + 1) * K * sizeof(float));
  c.total = off;
  return c;
}

}  // namespace

// Images are filtered kBatch at a time (grid.z = image); each has its own lattice region in the workspace.
constexpr int kBilateralBatch = 8;

extern "C" size_t acr_bilateral_workspace(int N, int K, int H, int W) {
  if (N <= 0 || K <= 0 || H <= 0 || W <= 0) return 0;
  return carve(nullptr, K, H, W).total * (size_t)(N < kBilateralBatch ? N : kBilateralBatch);
}

extern "C" int acr_bilateral_batch(const float* images, const float* ins, float* outs,
                                   int N, int K, int H, int W, float sigmargb, float sigmaxy,
                                   void* workspace, size_t workspace_bytes, int* lattice_size_host, void* stream) {
  ACR_REQUIRE(images && ins && outs && workspace, ACR_E_INVAL, "acr_bilateral_batch: null pointer");
  ACR_REQUIRE(N > 0 && K > 0 && H > 0 && W > 0, ACR_E_INVAL, "acr_bilateral_batch: bad shape");
  ACR_REQUIRE(sigmargb > 0.f && sigmaxy > 0.f, ACR_E_INVAL, "acr_bilateral_batch: sigma <= 0");
  ACR_REQUIRE((long long)H * W * D6 < (1ll << 28), ACR_E_INVAL, "acr_bilateral_batch: image too large");
  ACR_REQUIRE(((uintptr_t)workspace & 255) == 0, ACR_E_ALIGN, "acr_bilateral_batch: workspace not 256-byte aligned");
  const Carve c = carve(workspace, K, H, W);
  ACR_REQUIRE(workspace_bytes >= acr_bilateral_workspace(N, K, H, W), ACR_E_NOMEM, "acr_bilateral_batch: workspace too small (%zu < %zu)",
              workspace_bytes, acr_bilateral_workspace(N, K, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W, HWpad = (HW + 3) / 4 * 4, nc = HWpad * D6;

  Scales sc;
  const float inv_std_dev = (float)(std::sqrt(2.0 / 3.0) * (D5 + 1));          // permutohedral.cpp:168
  for (int i = 0; i < D5; ++i) sc.s[i] = (float)(1.0 / std::sqrt((double)((i + 2) * (i + 1))) * inv_std_dev);  // :170-171
  const float alpha = 1.0f / (1.0f + powf(2.f, -(float)D5));                   // :568

  const size_t smem = (size_t)kPixTile * (K + 1) * sizeof(float);
  ACR_REQUIRE(smem <= 48 * 1024, ACR_E_INVAL, "acr_bilateral_batch: K=%d too large", K);
  const size_t ws = c.total;
  for (int n0 = 0; n0 < N; n0 += kBilateralBatch) {
    const int nb = (N - n0 < kBilateralBatch) ? N - n0 : kBilateralBatch;
    const float* img = images + (size_t)n0 * 3 * HW;
    const float* in = ins + (size_t)n0 * K * HW;
    float* out = outs + (size_t)n0 * K * HW;
    for (int i = 0; i < nb; ++i) {
      ACR_CUDA(cudaMemsetAsync((char*)c.table + ws * i, 0xff, (size_t)c.cap * sizeof(int), st));
      ACR_CUDA(cudaMemsetAsync((char*)c.counter + ws * i, 0, sizeof(int), st));
    }
    lattice_embed_kernel<<<dim3((HWpad + 255) / 256, 1, nb), 256, 0, st>>>(img, H, W, HWpad, sigmargb, sigmaxy, sc, c.ckey, c.bary, ws);
    if (int e = acr::check_launch("lattice_embed_kernel")) return e;
    lattice_insert_kernel<<<dim3((nc + 255) / 256, 1, nb), 256, 0, st>>>(c.ckey, nc, c.table, c.cap - 1, c.rep, c.vid, c.vcand, c.counter, ws);
    if (int e = acr::check_launch("lattice_insert_kernel")) return e;
    lattice_neighbors_kernel<<<dim3((nc * D6 + 255) / 256, 1, nb), 256, 0, st>>>(c.ckey, c.table, c.cap - 1, c.vid, c.vcand, c.counter, nc, c.nbr, ws);
    if (int e = acr::check_launch("lattice_neighbors_kernel")) return e;
    lattice_offset_kernel<<<dim3((nc + 255) / 256, 1, nb), 256, 0, st>>>(c.rep, c.vid, nc, ws);
    if (int e = acr::check_launch("lattice_offset_kernel")) return e;
    zero_values_kernel<<<dim3(148, 1, nb), 256, 0, st>>>(c.val0, c.val1, c.counter, K, ws);
    if (int e = acr::check_launch("zero_values_kernel")) return e;
    lattice_splat_kernel<<<dim3((HW + kPixTile - 1) / kPixTile, 1, nb), 256, smem, st>>>(in, K, HW, c.rep, c.bary, c.val0, ws);
    if (int e = acr::check_launch("lattice_splat_kernel")) return e;
    float* cur = c.val0;
    float* nxt = c.val1;
    for (int j = 0; j < D6; ++j) {
      lattice_blur_kernel<<<dim3(148, 1, nb), 256, 0, st>>>(cur, nxt, c.nbr + (size_t)j * nc, c.counter, K, ws);
      if (int e = acr::check_launch("lattice_blur_kernel")) return e;
      float* t = cur; cur = nxt; nxt = t;
    }
    lattice_slice_kernel<<<dim3((HW + kPixTile - 1) / kPixTile, 1, nb), 256, smem, st>>>(cur, K, HW, c.rep, c.bary, alpha, out, ws);
    if (int e = acr::check_launch("lattice_slice_kernel")) return e;
    if (lattice_size_host) {
      for (int i = 0; i < nb; ++i)
        ACR_CUDA(cudaMemcpyAsync(lattice_size_host + n0 + i, (char*)c.counter + ws * i, sizeof(int), cudaMemcpyDeviceToHost, st));
      ACR_CUDA(cudaStreamSynchronize(st));
    }
  }
  return 0;
}

extern "C" void bilateralfilter_batch_b200(const float* images_host, int len_images, const float* ins_host, int len_ins,
                                           float* outs_host, int len_outs,
                                           int N, int K, int H, int W, float sigmargb, float sigmaxy) {
  const long long hw = (long long)H * W;
  if (!images_host || !ins_host || !outs_host || N <= 0 || K <= 0 || hw <= 0 ||
      (long long)len_images != 3 * hw * N || (long long)len_ins != K * hw * N || len_outs != len_ins) {
    acr::set_error("bilateralfilter_batch_b200: buffer lengths do not match N=%d K=%d H=%d W=%d", N, K, H, W);
    return;
  }
  const size_t ws_bytes = acr_bilateral_workspace(N, K, H, W);
  float *d_img = nullptr, *d_in = nullptr, *d_out = nullptr;
  void* d_ws = nullptr;
  cudaError_t e = cudaSuccess;
  cudaStream_t st = nullptr;
  auto fail = [&](const char* what) {
    acr::set_error("bilateralfilter_batch_b200: %s: %s", what, cudaGetErrorString(e));
    cudaFree(d_img); cudaFree(d_in); cudaFree(d_out); cudaFree(d_ws);
    if (st) cudaStreamDestroy(st);
  };
  if ((e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) { st = nullptr; return fail("cudaStreamCreate"); }
  if ((e = cudaMalloc(&d_img, (size_t)len_images * 4)) != cudaSuccess) return fail("cudaMalloc(images)");
  if ((e = cudaMalloc(&d_in, (size_t)len_ins * 4)) != cudaSuccess) return fail("cudaMalloc(ins)");
  if ((e = cudaMalloc(&d_out, (size_t)len_outs * 4)) != cudaSuccess) return fail("cudaMalloc(outs)");
  if ((e = cudaMalloc(&d_ws, ws_bytes)) != cudaSuccess) return fail("cudaMalloc(workspace)");
  if ((e = cudaMemcpyAsync(d_img, images_host, (size_t)len_images * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D images");
  if ((e = cudaMemcpyAsync(d_in, ins_host, (size_t)len_ins * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D ins");
  const int rc = acr_bilateral_batch(d_img, d_in, d_out, N, K, H, W, sigmargb, sigmaxy, d_ws, ws_bytes, nullptr, st);
  if (rc != 0) {
    cudaFree(d_img); cudaFree(d_in); cudaFree(d_out); cudaFree(d_ws); cudaStreamDestroy(st);
    return;   // error string already set
  }
  if ((e = cudaMemcpyAsync(outs_host, d_out, (size_t)len_outs * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return fail("D2H outs");
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail("synchronize");
  cudaFree(d_img); cudaFree(d_in); cudaFree(d_out); cudaFree(d_ws);
  cudaStreamDestroy(st);
  acr::set_error("");
}
