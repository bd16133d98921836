// Throughput microbenchmark: MUFU ex2 on f32 vs packed bf16x2 (is the packed form 2 exps per MUFU op on sm_100a?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ex2_bench ex2_bench.cu && ./ex2_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  uint32_t b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-4f; b[i] = __float_as_uint(a[i]) >> 3; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.0f;          // 1 MUFU + 1 FADD per exp
      else b[i] = ex2bf2(b[i]) ^ 0x00010001u;            // 1 MUFU(?) + 1 LOP per 2 exps
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(b[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 4, threads = 512;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, threads>>>(out, iters, 1.0f); else k<1><<<blocks, threads>>>(out, iters, 1.0f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double instr = double(blocks) * threads * iters * 8;
      double exps = instr * (mode == 0 ? 1 : 2);
      printf("mode %s: %.3f ms, %.1f G ex2-instr/s, %.1f G exps/s (per SM per clk @1.9GHz: %.2f exps)\n", mode == 0 ? "f32" : "bf16x2", ms,
             instr / ms / 1e6, exps / ms / 1e6, exps / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
