// TMEM -> register read bandwidth per SM (tcgen05.ld 32x32b.x32), 1 or 2 CTAs per SM, 4 or 8 warps per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sliders_b200/csrc -o tmem_ld_bench tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "ptx.cuh"
using namespace sb200;

template <int PASSES>
__global__ void k(uint32_t* out, int iters, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  const int q = warp & 3;
  const int half = warp >> 2;  // with 8 warps: each pair splits the 128 columns
  const int ncols = blockDim.x == 256 ? 64 : 128;
  const uint32_t tl = tb + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int ps = 0; ps < PASSES; ++ps) {
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        tmem_ld_x32(tl + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

int main() {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 2 * 256 * 4); cudaMalloc(&cyc, 148 * 2 * 8);
  const int iters = 2000;
  for (int threads : {128, 256}) for (int ctas : {1, 2}) {
    k<1><<<148 * ctas, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, cyc, 148 * ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148 * ctas; ++i) avg += h[i]; avg /= 148 * ctas;
    const double bytes_per_cta = double(iters) * 128 * 128 * 4;  // one 128x128 fp32 tile per iteration
    printf("threads/CTA %d, CTAs/SM %d: %.0f cycles for %d tile reads -> %.1f B/clk per CTA, %.1f B/clk per SM (%s)\n",
           threads, ctas, avg, iters, bytes_per_cta / avg, ctas * bytes_per_cta / avg, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
