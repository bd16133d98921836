// Per-instruction cost of tcgen05.mma (M = 128, K = 16, bf16) as a function of N, SS vs TS operand mode, one or two
// CTAs per SM, issued back-to-back by one thread with everything else out of the way (operands: zeroed smem / TMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../sliders_b200/csrc -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "ptx.cuh"
using namespace sb200;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// mode 0: SS, one accumulator; 1: SS, two accumulators alternating; 2: TS (A from TMEM columns 448..479)
__global__ void __launch_bounds__(128) k(int n_cols, int mode, int iters, int tmem_cols, long long* cyc, long long* ns) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 64;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, tmem_cols); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(sm + 16384 + 32768 + 64);
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n_cols, 0);
    const uint32_t a_tmem = tb + (tmem_cols - 64);
    uint32_t phase = 0;
    const long long c0 = clock64();
    const unsigned long long g0 = gtime();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = umma_desc_sw128(sA + kk * 32), bd = umma_desc_sw128(sB + kk * 32);
        const uint32_t d = tb + ((mode == 1 && (kk & 1)) ? static_cast<uint32_t>(n_cols) : 0u);
        if (mode == 2) umma_ts(d, a_tmem + kk * 8, bd, idesc, 1); else umma_ss(d, ad, bd, idesc, 1);
      }
      if ((it & 15) == 15) {  // keep the queue bounded: wait for completion every 64 instructions
        umma_commit(bar);
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
    }
    umma_commit(bar);
    mbar_wait(bar, phase);
    cyc[blockIdx.x] = clock64() - c0;
    ns[blockIdx.x] = static_cast<long long>(gtime() - g0);
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, tmem_cols); }
}

int main() {
  long long *cyc, *ns; cudaMalloc(&cyc, 296 * 8); cudaMalloc(&ns, 296 * 8);
  const int smem = 16384 + 32768 + 1024 + 1024, iters = 4096;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"SS one acc", "SS two acc", "TS (A in TMEM)"};
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int mode = 0; mode < 3; ++mode)
      for (int n : {64, 128, 256}) {
        const int tmem_cols = ctas == 2 ? 256 : 512;
        if (mode == 1 && 2 * n > tmem_cols - 64) continue;
        if (n > tmem_cols - 64 && mode == 2) continue;
        k<<<148 * ctas, 128, smem>>>(n, mode, iters, tmem_cols, cyc, ns);
        cudaError_t e = cudaDeviceSynchronize();
        long long hc[296], hn[296];
        cudaMemcpy(hc, cyc, 148 * ctas * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hn, ns, 148 * ctas * 8, cudaMemcpyDeviceToHost);
        double c = 0, t = 0; for (int i = 0; i < 148 * ctas; ++i) { c += hc[i]; t += hn[i]; }
        c /= 148 * ctas; t /= 148 * ctas;
        const double per = c / (4.0 * iters), per_ns = t / (4.0 * iters);
        printf("CTAs/SM %d  %-15s N=%3d: %6.1f clk (%5.1f ns) per UMMA per CTA -> %6.1f clk per UMMA per SM; ideal math %3d clk; "
               "%.0f TFLOP/s chip (%s)\n", ctas, names[mode], n, per, per_ns, per / ctas, n / 2,
               148.0 * ctas * 2.0 * 128 * n * 16 / per_ns / 1e3, cudaGetErrorString(e));
      }
  return 0;
}
