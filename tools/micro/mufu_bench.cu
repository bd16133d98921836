// MUFU.EX2 issue rate per SM sub-partition, alone and in the softmax instruction mix, for 1..8 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench tools/micro/mufu_bench.cu && ./mufu_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MODE>
__global__ void k(float* out, long long* clk, int iters, float scale, float m) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float p0, p1, p2, p3;
      if (MODE == 0) {  // MUFU only (independent)
        p0 = ex2(v[i]); p1 = ex2(v[i + 1]); p2 = ex2(v[i + 2]); p3 = ex2(v[i + 3]);
        v[i] = p0; v[i + 1] = p1; v[i + 2] = p2; v[i + 3] = p3;
      } else {          // softmax mix: FFMA, MUFU, FADD, pack
        p0 = ex2(fmaf(v[i], scale, -m)); p1 = ex2(fmaf(v[i + 1], scale, -m));
        p2 = ex2(fmaf(v[i + 2], scale, -m)); p3 = ex2(fmaf(v[i + 3], scale, -m));
        s0 += p0; s1 += p1; s2 += p2; s3 += p3;
        uint32_t a, b;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(p1), "f"(p0));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(p3), "f"(p2));
        acc ^= a + b;
        v[i] = -p0; v[i + 1] = -p1; v[i + 2] = -p2; v[i + 3] = -p3;
      }
    }
  }
  const long long t1 = clock64();
  float r = s0 + s1 + s2 + s3 + __uint_as_float(acc);
#pragma unroll
  for (int i = 0; i < 32; ++i) r += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int wps : {1, 2, 4, 8}) {
      const int threads = wps * 4 * 32;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, clk, iters, 0.125f, 1.f);
        else k<1><<<148, threads>>>(out, clk, iters, 0.125f, 1.f);
      }
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += h[i];
      avg /= 148;
      const double mufu_per_subpart = double(iters) * 32 * wps;  // warp-level MUFU instructions per sub-partition
      printf("%s  %d warps/sub-partition: %.2f clk per warp-MUFU per sub-partition (%.1f lanes/clk/SM)\n",
             mode ? "softmax mix" : "MUFU only  ", wps, avg / mufu_per_subpart, 128.0 * mufu_per_subpart / avg);
    }
  return 0;
}
