// Microbenchmark: TMEM -> register read bandwidth of tcgen05.ld.32x32b.x32 with 4 / 8 / 16 reading warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../acr_wsss_b200/csrc/tc_common.cuh"

__global__ void __launch_bounds__(640) tmem_read_kernel(int iters, int nread_warps, unsigned long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base;
  float acc = 0.f;
  unsigned long long t0 = 0, t1 = 0;
  if (warp >= 4 && warp < 4 + nread_warps) {
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int grp = (warp - 4) >> 2;                 // column group of 128
    uint32_t r[32];
    asm volatile("bar.sync 1, %0;" ::"r"(nread_warps * 32) : "memory");
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::tmem_ld32(tmem + lane_off + ((grp * 128 + c * 32) & 511), r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
      }
    }
    t1 = clock64();
    if (threadIdx.x == 128) cycles[blockIdx.x] = t1 - t0;
  }
  if (acc == 123.456f) sink[0] = acc;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tmem); }
}

int main() {
  unsigned long long* d_c; float* d_s;
  cudaMalloc(&d_c, 148 * 8); cudaMalloc(&d_s, 4);
  for (int nw : {4, 8, 16}) {
    const int iters = 2000;
    tmem_read_kernel<<<148, 640>>>(iters, nw, d_c, d_s);
    cudaDeviceSynchronize();
    unsigned long long c[148]; cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
    // bytes read per SM: each reading warp reads 32 lanes x 128 cols x 4 B per iteration
    double bytes = (double)nw * 32 * 128 * 4 * iters;
    printf("read warps %2d: %llu cycles, %.1f B/clk/SM  (err=%s)\n", nw, c[0], bytes / (double)c[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
