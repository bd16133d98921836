// How fast can ONE warp feed the tensor pipe?  Round 1 measured a "floor" of ~176 SM cycles per tcgen05.mma whatever
// its N (tools/micro/umma_bench.cu) and read it as hardware.  The ncu source page of the CTA-pair GEMM shows the MMA
// thread busy on its own instruction stream instead (ELECT / R2UR.BROADCAST / PLOP3 / BRA.U.ANY retry loop that ptxas
// wraps around every uniform-datapath instruction issued from a `lane == 0` branch, ~20 dependent fixed-latency
// instructions per UTCHMMA), so this benchmark times the issue LOOP FORMS themselves, with zeroed operands, no TMA and
// no epilogue:
//   form 0  lane-0 branch, one thread runs the loop (round-1 production form)
//   form 1  lane 0 polls the barrier, __syncwarp, elect.sync region issues (round-1 "elect" form)
//   form 2  converged warp: every lane polls the barrier, one elect.sync region issues the 4 MMAs + commit
//   form 3  form 2 with the barrier handshake removed (pure issue rate; commit to a barrier nobody waits on)
//   form 4  form 0 with the handshake removed (reproduces umma_bench)
// handshake = a second warp plays the TMA producer: waits empty[s], arrives full[s] (S = 4 stages, 4 MMAs per stage).
// cta_group 2 = cluster of two CTAs, leader issues M = 256, commits multicast to both CTAs (like gemm_kernel<2>).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../../sliders_b200/csrc -o issue_bench issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "ptx.cuh"
using namespace sb200;

constexpr int kS = 4;

__device__ __forceinline__ void mbar_wait_all(uint32_t bar, uint32_t parity) {
  // converged polling: every lane spins on try_wait (no clock64 watchdog in the loop body)
  while (!mbar_try_wait(bar, parity)) {
  }
}

template <int kCtas, int kForm>
__global__ void __launch_bounds__(128) k(int n_cols, int kblocks, long long* cyc, long long* ns) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t sA = base, sB = base + 16384, bars = base + 16384 + 32768;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kS, bar_done = bars + 16 * kS, slot = bars + 16 * kS + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kS; ++i) {
      mbar_init(bar_full + 8 * i, kCtas);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kCtas == 2) { tmem_alloc_2cta(slot, 512); tmem_relinquish_2cta(); }
    else { tmem_alloc(slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(sm + 16384 + 32768 + 16 * kS + 16);
  constexpr bool handshake = kForm <= 2;

  if (warp == 1 && handshake) {
    // stand-in for the TMA producer (every CTA): wait for the stage to be free, mark it full on the leader
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if constexpr (kCtas == 2) mbar_arrive_cluster(mapa_u32(bar_full + 8 * stage, 0));
        else mbar_arrive(bar_full + 8 * stage);
        if (++stage == kS) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(128 * kCtas, n_cols, 0);
    const uint32_t a_lo = (sA & 0x3FFFF) >> 4, b_lo = (sB & 0x3FFFF) >> 4;
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long c0 = 0; unsigned long long g0 = 0;
    auto issue4 = [&](int kb, int stage) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = (static_cast<uint64_t>(kHi) << 32) | (a_lo + 2u * kk);
        const uint64_t bd = (static_cast<uint64_t>(kHi) << 32) | (b_lo + 2u * kk);
        const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
        if constexpr (kCtas == 2) umma_ss_2cta(tb, ad, bd, idesc, acc); else umma_ss(tb, ad, bd, idesc, acc);
      }
      if constexpr (kCtas == 2) umma_commit_2cta(bar_empty + 8 * stage, 3); else umma_commit(bar_empty + 8 * stage);
    };
    if constexpr (kForm == 0 || kForm == 4) {
      if (lane == 0) {
        c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          if (kForm == 0) mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          issue4(kb, stage);
          if (++stage == kS) { stage = 0; phase ^= 1u; }
        }
        if constexpr (kCtas == 2) umma_commit_2cta(bar_done, 1); else umma_commit(bar_done);
        mbar_wait(bar_done, 0);
      }
    } else {
      c0 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        if (kForm == 1) {
          if (lane == 0) mbar_wait(bar_full + 8 * stage, phase);
          __syncwarp();
        } else if (kForm == 2) {
          mbar_wait_all(bar_full + 8 * stage, phase);
        }
        tc_fence_after();
        if (elect_one()) issue4(kb, stage);
        if (kForm == 1) __syncwarp();
        if (++stage == kS) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) {
        if constexpr (kCtas == 2) umma_commit_2cta(bar_done, 1); else umma_commit(bar_done);
      }
      __syncwarp();
      mbar_wait_all(bar_done, 0);
    }
    if (lane == 0) {
      unsigned long long g1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
      cyc[blockIdx.x] = clock64() - c0;
      ns[blockIdx.x] = static_cast<long long>(g1 - g0);
    }
  }
  __syncwarp();
  tc_fence_before();
  if constexpr (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kCtas == 2) tmem_dealloc_2cta(tb, 512); else tmem_dealloc(tb, 512);
  }
}

template <int kCtas, int kForm>
static void run(int n, long long* cyc, long long* ns) {
  const int smem = 16384 + 32768 + 1024 + 1024, kblocks = 4096, grid = 148;
  cudaFuncSetAttribute(k<kCtas, kForm>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(cyc, 0, 296 * 8);
  cudaMemset(ns, 0, 296 * 8);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = kCtas;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k<kCtas, kForm>, n, kblocks, cyc, ns);
  cudaError_t e = cudaDeviceSynchronize();
  long long hc[148], hn[148];
  cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hn, ns, grid * 8, cudaMemcpyDeviceToHost);
  double c = 0, t = 0;
  int cnt = 0;
  for (int i = 0; i < grid; ++i)
    if (hc[i] > 0) { c += hc[i]; t += hn[i]; ++cnt; }
  c /= cnt; t /= cnt;
  const double per = c / (4.0 * kblocks), per_ns = t / (4.0 * kblocks);
  const char* names[5] = {"lane0 + handshake", "lane0-poll/elect + handshake", "converged/elect + handshake",
                          "converged/elect, no handshake", "lane0, no handshake"};
  printf("cta_group %d  %-30s N=%3d: %6.1f clk (%5.1f ns) per UMMA; ideal math %3d clk; %5.0f TFLOP/s chip (%s)\n", kCtas,
         names[kForm], n, per, per_ns, n / 2, 148.0 * 2.0 * 128 * n * 16 / per_ns / 1e3, cudaGetErrorString(e));
}

int main() {
  long long *cyc, *ns;
  cudaMalloc(&cyc, 296 * 8);
  cudaMalloc(&ns, 296 * 8);
  for (int n : {64, 128, 256}) {
    run<1, 4>(n, cyc, ns);
    run<1, 3>(n, cyc, ns);
    run<1, 0>(n, cyc, ns);
    run<1, 1>(n, cyc, ns);
    run<1, 2>(n, cyc, ns);
    run<2, 4>(n, cyc, ns);
    run<2, 3>(n, cyc, ns);
    run<2, 0>(n, cyc, ns);
    run<2, 1>(n, cyc, ns);
    run<2, 2>(n, cyc, ns);
  }
  return 0;
}
