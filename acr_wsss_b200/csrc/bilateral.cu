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
// Eleven launches per batch of up to 8 images (grid.z = image); round 1 needed 13 launches + 16 memsets and 450 us at N = 8, K = 21:
//   embed      per pixel: 6 candidate keys + barycentric weights, written plane-major (candidate c = r*HWp + pixel, so every
//              store is coalesced); the same threads clear the hash table and the vertex counter
//   insert     per candidate; neighbouring pixels mostly fall on the same lattice vertex, so a warp first groups its lanes by
//              key (__match_any_sync on a 32-bit hash, verified against the group leader's key) and only leaders probe the table
//   prepare    one persistent launch for three independent jobs: blur neighbours of every vertex (2 table lookups per
//              (vertex, axis)), candidate -> vertex id + 1, zeroing of the value arrays
//   splat      a CTA sorts the 768 (pixel, remainder) pairs of its 16 x 8 pixel block by vertex in shared memory (integer atomics
//              only) and issues ONE red.global.add.v4.f32 per (vertex, 4 planes) (K padded to a multiple of 4: 16-byte value rows)
//   blur       d+1 passes, one launch each, (vertex, 4 planes) items with 4 independent gather chains per thread
//   slice      per (pixel, 4 planes): six 128-bit gathers, same accumulation order as the reference
//
// Data layout in HBM (per image, carved from the caller's workspace; nc = 6*HWp candidates, HWp = HW rounded up to 4):
//   ka,kb,kc [nc] u32      packed int16 key of every candidate (k0|k1<<16, k2|k3<<16, k4)
//   bary  [nc] f32         barycentric weight of the candidate
//   table [cap]  i32       open-addressing hash: representative candidate index or -1 (cap = pow2 >= 1.25*nc)
//   rep   [nc] i32         candidate -> representative candidate; `offs` = vertex id + 1 is a separate array
//   vid   [nc] i32         representative candidate -> dense vertex id
//   vcand [nc] i32         vertex id -> representative candidate
//   nbr   [6][nc] int2     blur neighbours (vertex id + 1; 0 = missing)
//   val0/val1 [(nc+1) * Kp] f32   lattice values, vertex-major with the Kp = roundup(K, 4) planes contiguous; slot 0 = zeros
#include "common.cuh"
#include <cmath>
#include <mutex>
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

// per-image arrays (pointers of image 0; image n lives `ws` bytes further)
struct Lattice {
  uint32_t *ka, *kb, *kc;
  float* bary;
  int *table, *rep, *offs, *vid, *vcand, *counter;
  int2* nbr;
  float *val0, *val1;
  uint32_t cap;
  size_t ws;
  int nc, Kp;
};
__device__ __forceinline__ Lattice for_image(const Lattice& l, int n) {
  Lattice r = l;
  r.ka = img_ptr(l.ka, l.ws, n); r.kb = img_ptr(l.kb, l.ws, n); r.kc = img_ptr(l.kc, l.ws, n);
  r.bary = img_ptr(l.bary, l.ws, n);
  r.table = img_ptr(l.table, l.ws, n); r.rep = img_ptr(l.rep, l.ws, n); r.offs = img_ptr(l.offs, l.ws, n);
  r.vid = img_ptr(l.vid, l.ws, n); r.vcand = img_ptr(l.vcand, l.ws, n); r.counter = img_ptr(l.counter, l.ws, n);
  r.nbr = img_ptr(l.nbr, l.ws, n);
  r.val0 = img_ptr(l.val0, l.ws, n); r.val1 = img_ptr(l.val1, l.ws, n);
  return r;
}

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
// 32-bit mix of the packed key (the reference's multiplicative hash, permutohedral.cpp:42-49, costs five 64-bit multiplies per
// probe; the hash function is free to choose -- results are indexed by key)
__device__ __forceinline__ uint32_t key_hash(const Key& key) {
  uint32_t h = key.a * 0x9E3779B1u;
  h = (h ^ (h >> 15)) + key.b * 0x85EBCA77u;
  h = (h ^ (h >> 13)) + key.c * 0xC2B2AE3Du;
  h ^= h >> 16;
  return h * 0x27D4EB2Fu;
}
__device__ __forceinline__ Key load_key(const Lattice& l, int c) { return Key{l.ka[c], l.kb[c], l.kc[c]}; }

// Step 1: per pixel, lattice coordinates -> 6 candidate keys + barycentric weights (plane-major: candidate r*HWp + pixel).
__global__ void __launch_bounds__(256)
lattice_embed_kernel(const float* __restrict__ image, int H, int W, int HWpad, float sigmargb, float sigmaxy, Scales sc, Lattice lat0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  const Lattice lat = for_image(lat0, blockIdx.z);
  // the hash table and the vertex counter of this image are cleared by the same launch
  for (uint32_t e = idx; e < lat.cap; e += gridDim.x * blockDim.x) lat.table[e] = -1;
  if (idx == 0) *lat.counter = 0;
  if (idx >= HWpad) return;
  image += (size_t)blockIdx.z * 3 * HW;
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
    const Key k = pack_key(key);
    const size_t c = (size_t)r * HWpad + idx;
    lat.ka[c] = k.a; lat.kb[c] = k.b; lat.kc[c] = k.c;
    lat.bary[c] = bary[r];
  }
}

// Step 2: insert every candidate; the CAS winner of a slot becomes the representative and draws a dense id.
__device__ __forceinline__ int lattice_insert(const Lattice& lat, const Key& k, uint32_t h, int c) {
  const uint32_t mask = lat.cap - 1;
  h &= mask;
  while (true) {
    int e = ((volatile int*)lat.table)[h];
    if (e == -1) {
      e = atomicCAS(&lat.table[h], -1, c);
      if (e == -1) {
        const int id = atomicAdd(lat.counter, 1);
        lat.vid[c] = id;
        lat.vcand[id] = c;
        return c;
      }
    }
    if (key_eq(load_key(lat, e), k)) return e;
    h = (h + 1) & mask;
  }
}

__global__ void __launch_bounds__(256)
lattice_insert_kernel(Lattice lat0) {
  const Lattice lat = for_image(lat0, blockIdx.z);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = c < lat.nc;
  Key k{0u, 0u, 0u};
  uint32_t h = 0x80000000u | (uint32_t)lane;          // lanes past the end: a group of their own, never inserted
  if (valid) { k = load_key(lat, c); h = key_hash(k) & 0x7fffffffu; }
  // lanes of a warp are consecutive pixels of one remainder plane: most of them share a vertex.  Group by hash, check the key
  // against the group leader's; only leaders (and the rare hash-equal, key-different lane) walk the table.
  const unsigned peers = __match_any_sync(0xffffffffu, h);
  const int leader = __ffs(peers) - 1;
  const Key kl{__shfl_sync(0xffffffffu, k.a, leader), __shfl_sync(0xffffffffu, k.b, leader), __shfl_sync(0xffffffffu, k.c, leader)};
  const bool own = valid && (lane == leader || !key_eq(k, kl));
  int res = -1;
  if (own) res = lattice_insert(lat, k, h, c);
  __syncwarp();
  const int lead_res = __shfl_sync(0xffffffffu, res, leader);
  if (valid) lat.rep[c] = own ? res : lead_res;
}

__device__ __forceinline__ int lattice_find(const Lattice& lat, const short* ks) {
  const Key k = pack_key(ks);
  const uint32_t mask = lat.cap - 1;
  uint32_t h = key_hash(k) & 0x7fffffffu & mask;
  while (true) {
    const int e = lat.table[h];
    if (e == -1) return 0;
    if (key_eq(load_key(lat, e), k)) return lat.vid[e] + 1;
    h = (h + 1) & mask;
  }
}

// Step 3 (one persistent launch, three independent jobs): blur neighbours of every vertex along each of the d+1 axes,
// candidate -> vertex id + 1, zero-fill of the value arrays (val0 entirely, slot 0 of val1).
__global__ void __launch_bounds__(256)
lattice_prepare_kernel(Lattice lat0) {
  const Lattice lat = for_image(lat0, blockIdx.z);
  const int M = *lat.counter;
  const int nthreads = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  int* nbr1 = reinterpret_cast<int*>(lat.nbr);
  for (int t = t0; t < M * D6 * 2; t += nthreads) {      // one table lookup per thread step: (vertex, axis, side)
    const int side = t & 1, vj = t >> 1;
    const int v = vj / D6, j = vj - v * D6;
    short key[D5], nk[D5];
    unpack_key(load_key(lat, lat.vcand[v]), key);
#pragma unroll
    for (int k = 0; k < D5; ++k) nk[k] = side ? key[k] + 1 : key[k] - 1;
#pragma unroll
    for (int k = 0; k < D5; ++k)
      if (k == j) nk[k] = side ? key[k] - D5 : key[k] + D5;
    nbr1[2 * ((size_t)j * lat.nc + v) + side] = lattice_find(lat, nk);
  }
  for (int c = t0; c < lat.nc; c += nthreads) lat.offs[c] = lat.vid[lat.rep[c]] + 1;
  const long long nz = ((long long)M + 1) * lat.Kp / 4;
  float4* z0 = reinterpret_cast<float4*>(lat.val0);
  for (long long i = t0; i < nz; i += nthreads) z0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t0 < lat.Kp) lat.val1[t0] = 0.f;
}

__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int kPixTile = 128;      // pixels per CTA (splat: 16 x 8 block of the image; slice: 128 consecutive pixels)
constexpr int kTileW = 16, kTileH = 8;
constexpr int kSlotProbes = 32;

// Splat.  A CTA owns a 16 x 8 block of pixels: its 768 (pixel, remainder) pairs fall on ~100-200 distinct lattice vertices.  Float
// atomics at L2 bounded the first version (50 M of them for N = 8, K = 21), and shared-memory float atomics are CAS loops (4x
// slower still, measured), so the pairs are SORTED by vertex inside the CTA with integer shared-memory atomics only:
//   1. input planes -> shared memory, transposed to tile4[pixel][quad] (float4 = 4 consecutive planes; odd row stride);
//   2. every pair finds its vertex's slot in a small open-addressing table (atomicCAS) and draws a position in the slot (atomicAdd);
//   3. exclusive scan of the slot counts, pairs scattered to order[] (counting sort);
//   4. one work item per (occupied slot, quad) sums its pairs from the tile and issues ONE red.global.add.v4.f32.
// Pairs that find no slot within kSlotProbes probes add to global memory directly (rare).
constexpr int kSplatSlots = 1024;    // >= the 768 pairs of a tile, so the table never fills up; 4 slots per thread in the scan
__global__ void __launch_bounds__(256)
lattice_splat_kernel(const float* __restrict__ in, int K, int H, int W, int HWpad, Lattice lat0) {
  extern __shared__ float4 smem4[];                    // tile4[kPixTile][Qs], then the int / float arrays below
  __shared__ int skey[kSplatSlots], scount[kSplatSlots], sstart[kSplatSlots], socc[kSplatSlots], warp_tot[8], warp_occ[8], n_occ;
  const Lattice lat = for_image(lat0, blockIdx.z);
  const int HW = H * W, Kp = lat.Kp, Q = Kp / 4;
  in += (size_t)blockIdx.z * K * HW;
  const int Qs = Q | 1;                                 // odd row stride (in float4): conflict-free 128-bit stores
  int* slot = reinterpret_cast<int*>(smem4 + (size_t)Qs * kPixTile);      // [D6 * kPixTile]: slot of the pair, -1 = none
  int* pos = slot + D6 * kPixTile;                                        // position inside the slot
  int* order = pos + D6 * kPixTile;                                       // pair indices sorted by slot
  float* wts = reinterpret_cast<float*>(order + D6 * kPixTile);
  const int tiles_x = (W + kTileW - 1) / kTileW;
  const int x0 = (blockIdx.x % tiles_x) * kTileW, y0 = (blockIdx.x / tiles_x) * kTileH;
  for (int i = threadIdx.x; i < kSplatSlots; i += blockDim.x) { skey[i] = 0; scount[i] = 0; }
  for (int e = threadIdx.x; e < Q * kPixTile; e += blockDim.x) {      // four coalesced plane reads -> one 16-byte store
    const int q = e / kPixTile, p = e - q * kPixTile;
    const int x = x0 + (p % kTileW), y = y0 + (p / kTileW);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (4 * q + j < K && x < W && y < H) ? __ldg(in + (size_t)(4 * q + j) * HW + (size_t)y * W + x) : 0.f;
    smem4[p * Qs + q] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  for (int w = threadIdx.x; w < D6 * kPixTile; w += blockDim.x) {
    const int p = w % kPixTile, r = w / kPixTile;
    const int x = x0 + (p % kTileW), y = y0 + (p / kTileW);
    int sl = -1;
    float wt = 0.f;
    if (x < W && y < H) {
      const int pix = y * W + x;
      const int o = __ldg(lat.offs + (size_t)r * HWpad + pix);
      wt = __ldg(lat.bary + (size_t)r * HWpad + pix);
      const uint32_t h = ((uint32_t)o * 0x9E3779B1u) >> 7;
      for (int t = 0; t < kSlotProbes; ++t) {
        const int i = (int)((h + t) & (uint32_t)(kSplatSlots - 1));
        const int prev = atomicCAS(&skey[i], 0, o);
        if (prev == 0 || prev == o) { sl = i; break; }
      }
      if (sl >= 0) {
        pos[w] = atomicAdd(&scount[sl], 1);
      } else {                                          // table neighbourhood full (rare): this pair adds to global memory itself
        for (int q = 0; q < Q; ++q) {
          const float4 x = smem4[p * Qs + q];
          red_add_v4(lat.val0 + (size_t)o * Kp + 4 * q, make_float4(wt * x.x, wt * x.y, wt * x.z, wt * x.w));
        }
      }
    }
    slot[w] = sl;
    wts[w] = wt;
  }
  __syncthreads();
  {   // exclusive scan of scount (4 consecutive slots per thread) and compaction of the occupied slots
    constexpr int PER = kSplatSlots / 256;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    int c[PER], tot = 0, occ = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) { c[u] = scount[threadIdx.x * PER + u]; tot += c[u]; occ += c[u] > 0; }
    int v = tot, vo = occ;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o), no = __shfl_up_sync(0xffffffffu, vo, o);
      if (lane >= o) { v += n; vo += no; }
    }
    if (lane == 31) { warp_tot[wp] = v; warp_occ[wp] = vo; }
    __syncthreads();
    int base = v - tot, obase = vo - occ;
    for (int i = 0; i < wp; ++i) { base += warp_tot[i]; obase += warp_occ[i]; }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      sstart[threadIdx.x * PER + u] = base;
      base += c[u];
      if (c[u] > 0) socc[obase++] = threadIdx.x * PER + u;
    }
    if (threadIdx.x == blockDim.x - 1) n_occ = obase;
  }
  __syncthreads();
  for (int w = threadIdx.x; w < D6 * kPixTile; w += blockDim.x) {
    const int sl = slot[w];
    if (sl >= 0) order[sstart[sl] + pos[w]] = w;
  }
  __syncthreads();
  const int nitems = n_occ * Q;
  for (int w = threadIdx.x; w < nitems; w += blockDim.x) {
    const int sidx = socc[w / Q], q = w % Q;
    const int beg = sstart[sidx], n = scount[sidx];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      const int pw = order[beg + j];
      const float wt = wts[pw];
      const float4 x = smem4[(pw % kPixTile) * Qs + q];
      acc.x = fmaf(wt, x.x, acc.x); acc.y = fmaf(wt, x.y, acc.y); acc.z = fmaf(wt, x.z, acc.z); acc.w = fmaf(wt, x.w, acc.w);
    }
    red_add_v4(lat.val0 + (size_t)skey[sidx] * Kp + 4 * q, acc);
  }
}

// One blur pass along axis j: new = old + 0.5 * (old[n1] + old[n2]).  Work item = (vertex, 4 planes); the chain neighbour ids ->
// three gathers is pure L2 latency, so every thread keeps kBlurIlp independent items in flight.  One launch per pass over the whole
// GPU: a single launch with one 8-CTA cluster per image and cluster barriers between the passes was measured as well -- no faster
// on small lattices (75 vs 78 us at 15 k vertices) and 64 of 148 SMs wide on large ones (noise images: 300 k vertices per image).
constexpr int kBlurIlp = 4;
__global__ void __launch_bounds__(256)
lattice_blur_kernel(const float* __restrict__ old0, float* __restrict__ new0, int j, Lattice lat0) {
  const Lattice lat = for_image(lat0, blockIdx.z);
  const float* oldv = img_ptr(old0, lat.ws, blockIdx.z);
  float* newv = img_ptr(new0, lat.ws, blockIdx.z);
  const long long M = *lat.counter;
  const int Q = lat.Kp / 4;
  const long long total = M * Q;
  const int nthreads = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int2* nbr = lat.nbr + (size_t)j * lat.nc;
  for (long long t = t0; t < total; t += (long long)kBlurIlp * nthreads) {
    size_t self[kBlurIlp], n1[kBlurIlp], n2[kBlurIlp];
    bool ok[kBlurIlp];
#pragma unroll
    for (int u = 0; u < kBlurIlp; ++u) {
      const long long tt = t + (long long)u * nthreads;
      ok[u] = tt < total;
      const int v = ok[u] ? (int)(tt / Q) : 0, q = ok[u] ? (int)(tt - (long long)v * Q) : 0;
      const int2 nb = __ldg(nbr + v);
      self[u] = (size_t)(v + 1) * lat.Kp + 4 * q;
      n1[u] = (size_t)nb.x * lat.Kp + 4 * q;
      n2[u] = (size_t)nb.y * lat.Kp + 4 * q;
    }
    float4 a[kBlurIlp], b[kBlurIlp], c[kBlurIlp];
#pragma unroll
    for (int u = 0; u < kBlurIlp; ++u) {
      a[u] = __ldg(reinterpret_cast<const float4*>(oldv + n1[u]));
      b[u] = __ldg(reinterpret_cast<const float4*>(oldv + n2[u]));
      c[u] = __ldg(reinterpret_cast<const float4*>(oldv + self[u]));
    }
#pragma unroll
    for (int u = 0; u < kBlurIlp; ++u) {
      float4 o;
      o.x = __fadd_rn(c[u].x, __fmul_rn(0.5f, __fadd_rn(a[u].x, b[u].x)));
      o.y = __fadd_rn(c[u].y, __fmul_rn(0.5f, __fadd_rn(a[u].y, b[u].y)));
      o.z = __fadd_rn(c[u].z, __fmul_rn(0.5f, __fadd_rn(a[u].z, b[u].z)));
      o.w = __fadd_rn(c[u].w, __fmul_rn(0.5f, __fadd_rn(a[u].w, b[u].w)));
      if (ok[u]) *reinterpret_cast<float4*>(newv + self[u]) = o;
    }
  }
}

// Slice: out = alpha * sum_r bary_r * values[vertex_r], accumulated in the reference's order.  A group of G lanes (G = 8, 16 or 32
// >= Kp/4) serves one pixel: its six vertex ids / weights are loaded once per lane (broadcast inside the group) and every lane
// gathers 16 bytes of each vertex row, so a row is read with contiguous requests; neighbouring pixels share vertices (L1 hits).
__global__ void __launch_bounds__(256)
lattice_slice_kernel(const float* __restrict__ values_img0, int K, int HW, int HWpad, float alpha, float* __restrict__ out, Lattice lat0) {
  extern __shared__ float4 smem4[];   // tile4[Kp/4][kPixTile]
  const Lattice lat = for_image(lat0, blockIdx.z);
  const float* values = img_ptr(values_img0, lat.ws, blockIdx.z);
  out += (size_t)blockIdx.z * K * HW;
  const int p0 = blockIdx.x * kPixTile;
  const int Q = lat.Kp / 4;
  const int G = Q <= 8 ? 8 : (Q <= 16 ? 16 : 32), ppw = 32 / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane % G, lp = lane / G;
  constexpr int kPixPerWarp = kPixTile / 8;
  for (int qq = q; qq < Q; qq += G) {                  // (one trip unless K > 128)
    for (int it = 0; it < kPixPerWarp / ppw; ++it) {
      const int p = warp * kPixPerWarp + it * ppw + lp;
      const int pix = p0 + p;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix < HW) {
        int o[D6];
        float wt[D6];
#pragma unroll
        for (int r = 0; r < D6; ++r) {
          o[r] = __ldg(lat.offs + (size_t)r * HWpad + pix);
          wt[r] = __fmul_rn(__ldg(lat.bary + (size_t)r * HWpad + pix), alpha);
        }
        float4 v[D6];
#pragma unroll
        for (int r = 0; r < D6; ++r) v[r] = __ldg(reinterpret_cast<const float4*>(values + (size_t)o[r] * lat.Kp + 4 * qq));
#pragma unroll
        for (int r = 0; r < D6; ++r) {
          s.x = __fadd_rn(s.x, __fmul_rn(wt[r], v[r].x));
          s.y = __fadd_rn(s.y, __fmul_rn(wt[r], v[r].y));
          s.z = __fadd_rn(s.z, __fmul_rn(wt[r], v[r].z));
          s.w = __fadd_rn(s.w, __fmul_rn(wt[r], v[r].w));
        }
      }
      smem4[qq * kPixTile + p] = s;
    }
  }
  __syncthreads();
  const float* tile = reinterpret_cast<const float*>(smem4);
  for (int e = threadIdx.x; e < K * kPixTile; e += blockDim.x) {
    const int k = e / kPixTile, p = e - k * kPixTile;
    if (p0 + p < HW) out[(size_t)k * HW + p0 + p] = tile[((k >> 2) * kPixTile + p) * 4 + (k & 3)];
  }
}

Lattice carve(void* base, int K, int H, int W) {
  Lattice c{};
  const size_t HW = ((size_t)H * W + 3) / 4 * 4, nc = HW * D6;   // incl. the reference's phantom pad pixels
  uint32_t cap = 1;
  while (cap < nc + nc / 4) cap <<= 1;      // worst case (every candidate a distinct vertex): load factor <= 0.8
  c.cap = cap;
  c.nc = (int)nc;
  c.Kp = (K + 3) / 4 * 4;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += acr::align_up(bytes, 256); return (char*)base + o; };
  c.counter = (int*)take(256);
  c.ka = (uint32_t*)take(nc * sizeof(uint32_t));
  c.kb = (uint32_t*)take(nc * sizeof(uint32_t));
  c.kc = (uint32_t*)take(nc * sizeof(uint32_t));
  c.bary = (float*)take(nc * sizeof(float));
  c.table = (int*)take((size_t)cap * sizeof(int));
  c.rep = (int*)take(nc * sizeof(int));
  c.offs = (int*)take(nc * sizeof(int));
  c.vid = (int*)take(nc * sizeof(int));
  c.vcand = (int*)take(nc * sizeof(int));
  c.nbr = (int2*)take(nc * D6 * sizeof(int2));
  c.val0 = (float*)take((nc + 1) * c.Kp * sizeof(float));
  c.val1 = (float*)take((nc + 1) * c.Kp * sizeof(float));
  c.ws = off;
  return c;
}

}  // namespace

// Images are filtered kBatch at a time (grid.z = image); each has its own lattice region in the workspace.
constexpr int kBilateralBatch = 8;

extern "C" size_t acr_bilateral_workspace(int N, int K, int H, int W) {
  if (N <= 0 || K <= 0 || H <= 0 || W <= 0) return 0;
  return carve(nullptr, K, H, W).ws * (size_t)(N < kBilateralBatch ? N : kBilateralBatch);
}

extern "C" int acr_bilateral_batch(const float* images, const float* ins, float* outs,
                                   int N, int K, int H, int W, float sigmargb, float sigmaxy,
                                   void* workspace, size_t workspace_bytes, int* lattice_size_host, void* stream) {
  ACR_REQUIRE(images && ins && outs && workspace, ACR_E_INVAL, "acr_bilateral_batch: null pointer");
  ACR_REQUIRE(N > 0 && K > 0 && H > 0 && W > 0, ACR_E_INVAL, "acr_bilateral_batch: bad shape");
  ACR_REQUIRE(sigmargb > 0.f && sigmaxy > 0.f, ACR_E_INVAL, "acr_bilateral_batch: sigma <= 0");
  ACR_REQUIRE((long long)H * W * D6 < (1ll << 28), ACR_E_INVAL, "acr_bilateral_batch: image too large");
  ACR_REQUIRE(((uintptr_t)workspace & 255) == 0, ACR_E_ALIGN, "acr_bilateral_batch: workspace not 256-byte aligned");
  const Lattice c = carve(workspace, K, H, W);
  ACR_REQUIRE(workspace_bytes >= acr_bilateral_workspace(N, K, H, W), ACR_E_NOMEM, "acr_bilateral_batch: workspace too small (%zu < %zu)",
              workspace_bytes, acr_bilateral_workspace(N, K, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W, HWpad = (HW + 3) / 4 * 4, nc = HWpad * D6;

  Scales sc;
  const float inv_std_dev = (float)(std::sqrt(2.0 / 3.0) * (D5 + 1));          // permutohedral.cpp:168
  for (int i = 0; i < D5; ++i) sc.s[i] = (float)(1.0 / std::sqrt((double)((i + 2) * (i + 1))) * inv_std_dev);  // :170-171
  const float alpha = 1.0f / (1.0f + powf(2.f, -(float)D5));                   // :568

  const size_t smem = (size_t)kPixTile * c.Kp * sizeof(float);      // slice: output tile
  ACR_REQUIRE(smem <= 48 * 1024, ACR_E_INVAL, "acr_bilateral_batch: K=%d too large (<= %d planes)", K, 48 * 1024 / (kPixTile * 4));
  // splat: input tile + four per-pair arrays
  const size_t smem_splat = (size_t)kPixTile * ((c.Kp / 4) | 1) * 16 + (size_t)D6 * kPixTile * 4 * sizeof(int);
  {
    static bool attr_set[64] = {false};
    int dev = 0;
    ACR_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      ACR_CUDA(cudaFuncSetAttribute(lattice_splat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr_set[dev] = true;
    }
  }
  const int splat_tiles = ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  for (int n0 = 0; n0 < N; n0 += kBilateralBatch) {
    const int nb = (N - n0 < kBilateralBatch) ? N - n0 : kBilateralBatch;
    const float* img = images + (size_t)n0 * 3 * HW;
    const float* in = ins + (size_t)n0 * K * HW;
    float* out = outs + (size_t)n0 * K * HW;
    lattice_embed_kernel<<<dim3((HWpad + 255) / 256, 1, nb), 256, 0, st>>>(img, H, W, HWpad, sigmargb, sigmaxy, sc, c);
    if (int e = acr::check_launch("lattice_embed_kernel")) return e;
    lattice_insert_kernel<<<dim3((nc + 255) / 256, 1, nb), 256, 0, st>>>(c);
    if (int e = acr::check_launch("lattice_insert_kernel")) return e;
    lattice_prepare_kernel<<<dim3(148 * 4, 1, nb), 256, 0, st>>>(c);
    if (int e = acr::check_launch("lattice_prepare_kernel")) return e;
    lattice_splat_kernel<<<dim3(splat_tiles, 1, nb), 256, smem_splat, st>>>(in, K, H, W, HWpad, c);
    if (int e = acr::check_launch("lattice_splat_kernel")) return e;
    {
      float* cur = c.val0;
      float* nxt = c.val1;
      for (int j = 0; j < D6; ++j) {
        lattice_blur_kernel<<<dim3(148 * 2, 1, nb), 256, 0, st>>>(cur, nxt, j, c);
        if (int e = acr::check_launch("lattice_blur_kernel")) return e;
        float* t = cur; cur = nxt; nxt = t;
      }
    }
    // d+1 = 6 passes (an even number): the result is back in val0
    lattice_slice_kernel<<<dim3((HW + kPixTile - 1) / kPixTile, 1, nb), 256, smem, st>>>(c.val0, K, HW, HWpad, alpha, out, c);
    if (int e = acr::check_launch("lattice_slice_kernel")) return e;
    if (lattice_size_host) {
      for (int i = 0; i < nb; ++i)
        ACR_CUDA(cudaMemcpyAsync(lattice_size_host + n0 + i, (char*)c.counter + c.ws * i, sizeof(int), cudaMemcpyDeviceToHost, st));
      ACR_CUDA(cudaStreamSynchronize(st));
    }
  }
  return 0;
}

// SWIG-shaped host entry (bilateralfilter.cpp:42-55): HOST buffers in, HOST buffer out.  The device buffers, the lattice workspace
// and the stream are cached per device and only grow (the reference allocates and frees its lattice on every call).
namespace {
struct HostCache {
  float *img = nullptr, *in = nullptr, *out = nullptr;
  void* ws = nullptr;
  size_t img_b = 0, io_b = 0, ws_b = 0;
  cudaStream_t st = nullptr;
};
std::mutex g_host_mutex;
HostCache g_host_cache[64];

cudaError_t grow(void** p, size_t* have, size_t need) {
  if (*have >= need) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc(p, need);
  if (e == cudaSuccess) *have = need;
  return e;
}
}  // namespace

extern "C" void bilateralfilter_batch_b200(const float* images_host, int len_images, const float* ins_host, int len_ins,
                                           float* outs_host, int len_outs,
                                           int N, int K, int H, int W, float sigmargb, float sigmaxy) {
  const long long hw = (long long)H * W;
  if (!images_host || !ins_host || !outs_host || N <= 0 || K <= 0 || hw <= 0 ||
      (long long)len_images != 3 * hw * N || (long long)len_ins != K * hw * N || len_outs != len_ins) {
    acr::set_error("bilateralfilter_batch_b200: buffer lengths do not match N=%d K=%d H=%d W=%d", N, K, H, W);
    return;
  }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) { acr::set_error("bilateralfilter_batch_b200: cudaGetDevice: %s", cudaGetErrorString(e)); return; }
  std::lock_guard<std::mutex> lock(g_host_mutex);
  HostCache& hc = g_host_cache[dev];
  auto fail = [&](const char* what) { acr::set_error("bilateralfilter_batch_b200: %s: %s", what, cudaGetErrorString(e)); };
  if (!hc.st && (e = cudaStreamCreateWithFlags(&hc.st, cudaStreamNonBlocking)) != cudaSuccess) { hc.st = nullptr; return fail("cudaStreamCreate"); }
  const size_t ws_bytes = acr_bilateral_workspace(N, K, H, W);
  size_t in_b = hc.io_b, out_b = hc.io_b;
  if ((e = grow((void**)&hc.img, &hc.img_b, (size_t)len_images * 4)) != cudaSuccess) return fail("cudaMalloc(images)");
  if ((e = grow((void**)&hc.in, &in_b, (size_t)len_ins * 4)) != cudaSuccess) { hc.io_b = 0; return fail("cudaMalloc(ins)"); }
  if ((e = grow((void**)&hc.out, &out_b, (size_t)len_outs * 4)) != cudaSuccess) { hc.io_b = 0; return fail("cudaMalloc(outs)"); }
  hc.io_b = in_b < out_b ? in_b : out_b;
  if ((e = grow(&hc.ws, &hc.ws_b, ws_bytes)) != cudaSuccess) return fail("cudaMalloc(workspace)");
  if ((e = cudaMemcpyAsync(hc.img, images_host, (size_t)len_images * 4, cudaMemcpyHostToDevice, hc.st)) != cudaSuccess) return fail("H2D images");
  if ((e = cudaMemcpyAsync(hc.in, ins_host, (size_t)len_ins * 4, cudaMemcpyHostToDevice, hc.st)) != cudaSuccess) return fail("H2D ins");
  if (acr_bilateral_batch(hc.img, hc.in, hc.out, N, K, H, W, sigmargb, sigmaxy, hc.ws, hc.ws_b, nullptr, hc.st) != 0) return;   // error string already set
  if ((e = cudaMemcpyAsync(outs_host, hc.out, (size_t)len_outs * 4, cudaMemcpyDeviceToHost, hc.st)) != cudaSuccess) return fail("D2H outs");
  if ((e = cudaStreamSynchronize(hc.st)) != cudaSuccess) return fail("synchronize");
  acr::set_error("");
}
