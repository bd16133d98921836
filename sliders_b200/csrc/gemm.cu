// LoRA-fused bf16 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M, N] = epilogue( A[M, K] . W[N, K]^T  [+ scale * (A . down^T) . up^T] )
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   — streams 128x64 A tiles and bn x 64 W tiles (+ the rt x 64 LoRA-down
//                                 tile) through an mbarrier ring of shared-memory stages
//   warp 1      MMA issuer     — one thread issues tcgen05.mma (UMMA 128 x bn x 16, bf16 -> fp32) into a
//                                 double-buffered TMEM accumulator; the LoRA down-projection is a second
//                                 128 x rt x 16 UMMA on the same A tile into spare TMEM columns
//   warps 2..5  epilogue       — tcgen05.ld the accumulator (one row per thread), apply the rank-r LoRA
//                                 up-projection, bias / time-embedding row bias / GEGLU / residual, store bf16
//
// The A operand has three addressing modes:
//   plain   2-D [M, K], optionally split along K over two sources (skip-connection concat, conv_shortcut)
//   conv    4-D NHWC box per filter tap (implicit GEMM): the box start is shifted by (kh-1, kw-1) and TMA's
//           out-of-bounds zero fill provides the padding; two sources along C give the concat
//   conv/2  stride-2 conv: four parity-plane descriptors (even/odd rows x even/odd columns)
//
// Replaces (reference): every nn.Linear / nn.Conv2d leaf that diffusers' UNet2DConditionModel executes
// under trainscripts/textsliders/train_util.py:242-247 together with the LoRA hook
// trainscripts/textsliders/lora.py:108-112.
#include "common.h"
#include "ptx.cuh"

namespace sb200 {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBarRegion = 1024;

struct GemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmL;
  int M, N, K;
  int bn;           // accumulator tile width (UMMA N), multiple of 16
  int ncols_out;    // output columns per tile (bn, or bn/2 with GEGLU)
  int Nout;         // output columns overall (N, or N/2 with GEGLU)
  int num_m_tiles, num_n_tiles;
  int kblocks;      // K / 64 (conv: 9 * cb_total)
  int stages;
  int stage_bytes;  // bytes of one smem stage == expected TMA transaction bytes
  int a_mode;       // 0 plain, 1 conv3x3 stride 1, 2 conv3x3 stride 2
  int kb_split;     // k-blocks (per tap) that come from source 0
  int cb_total;     // k-blocks per tap
  int H, W;         // conv OUTPUT spatial dims
  int flags;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* rowbias;
  int rows_per_batch;
  const __nv_bfloat16* resid;
  int ldr;
  __nv_bfloat16* out;
  int ldo;
  const __nv_bfloat16* lora_up;
  int lora_r, lora_rt, lora_group_n;
  float lora_scale;
  const float* lora_scale_dev;
};

template <int R>
__device__ __forceinline__ void lora_apply(float* f, const float* t, const __nv_bfloat16* up_rows) {
  // up_rows: 16 consecutive rows of [N, R] bf16 (same address for the whole warp -> L1 broadcast)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if constexpr (R == 4) {
      uint2 u = __ldg(reinterpret_cast<const uint2*>(up_rows + i * 4));
      f[i] += t[0] * bf16_lo(u.x) + t[1] * bf16_hi(u.x) + t[2] * bf16_lo(u.y) + t[3] * bf16_hi(u.y);
    } else {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(up_rows + i * 8));
      f[i] += t[0] * bf16_lo(u.x) + t[1] * bf16_hi(u.x) + t[2] * bf16_lo(u.y) + t[3] * bf16_hi(u.y) +
              t[4] * bf16_lo(u.z) + t[5] * bf16_hi(u.z) + t[6] * bf16_lo(u.w) + t[7] * bf16_hi(u.w);
    }
  }
}

__global__ void __launch_bounds__(kGemmThreads, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t bar_full = base;                 // S barriers
  const uint32_t bar_empty = base + 8u * S;       // S barriers
  const uint32_t bar_tfull = base + 16u * S;      // 2 barriers
  const uint32_t bar_tempty = bar_tfull + 16u;    // 2 barriers
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 512);
  const uint32_t tiles = base + kBarRegion;
  const bool has_lora = (p.flags & SB200_EPI_LORA) != 0;
  const bool geglu = (p.flags & SB200_EPI_GEGLU) != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    if (has_lora) tma_prefetch_desc(&p.tmL);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(bar_full + 8u * i, 1);
      mbar_init(bar_empty + 8u * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8u * i, 1);
      mbar_init(bar_tempty + 8u * i, 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  const uint32_t a_bytes = kBM * 128;
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * 128;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t % p.num_m_tiles;
        const int nt = t / p.num_m_tiles;
        const int m0 = mt * kBM;
        int b0 = 0, h0 = 0, w0 = 0;
        if (p.a_mode != 0) {
          const int hw = p.H * p.W;
          b0 = m0 / hw;
          const int rem = m0 - b0 * hw;
          h0 = rem / p.W;
          w0 = rem - h0 * p.W;
        }
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t full = bar_full + 8u * stage;
          mbar_expect_tx(full, static_cast<uint32_t>(p.stage_bytes));
          const uint32_t sA = tiles + static_cast<uint32_t>(stage) * p.stage_bytes;
          const uint32_t sB = sA + a_bytes;
          if (p.a_mode == 0) {
            const int src = kb < p.kb_split ? 0 : 1;
            const int kc = (src ? kb - p.kb_split : kb) * kBK;
            tma_load_2d(sA, &p.tmA[src], full, kc, m0);
          } else {
            const int tap = kb / p.cb_total;
            const int cb = kb - tap * p.cb_total;
            const int kh = tap / 3;
            const int kw = tap - kh * 3;
            if (p.a_mode == 1) {
              const int src = cb < p.kb_split ? 0 : 1;
              const int c0 = (src ? cb - p.kb_split : cb) * kBK;
              tma_load_4d(sA, &p.tmA[src], full, c0, w0 + kw - 1, h0 + kh - 1, b0);
            } else {
              // input row 2*ho + kh - 1: kh=0 -> odd plane, row ho-1; kh=1 -> even plane, row ho;
              // kh=2 -> odd plane, row ho
              const int ph = (kh == 1) ? 0 : 1;
              const int pw = (kw == 1) ? 0 : 1;
              const int dh = (kh == 0) ? -1 : 0;
              const int dw = (kw == 0) ? -1 : 0;
              tma_load_4d(sA, &p.tmA[ph * 2 + pw], full, cb * kBK, w0 + dw, h0 + dh, b0);
            }
          }
          if (geglu) {
            const int half = p.bn >> 1;
            tma_load_2d(sB, &p.tmB, full, kb * kBK, nt * half);
            tma_load_2d(sB + static_cast<uint32_t>(half) * 128, &p.tmB, full, kb * kBK,
                        (p.N >> 1) + nt * half);
          } else {
            tma_load_2d(sB, &p.tmB, full, kb * kBK, nt * p.bn);
          }
          if (has_lora) tma_load_2d(sB + b_bytes, &p.tmL, full, kb * kBK, 0);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kBM, p.bn);
      const uint32_t idesc_l = umma_idesc_bf16(kBM, has_lora ? p.lora_rt : 16);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(bar_tempty + 8u * as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * 256u;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(bar_full + 8u * stage, phase);
          tc_fence_after();
          const uint32_t sA = tiles + static_cast<uint32_t>(stage) * p.stage_bytes;
          const uint32_t sB = sA + a_bytes;
          const uint32_t sL = sB + b_bytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = umma_desc_sw128(sA + k * 32);
            const uint64_t bdesc = umma_desc_sw128(sB + k * 32);
            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
            umma_ss(d_tmem, adesc, bdesc, idesc, acc);
            if (has_lora) {
              const uint64_t ldesc = umma_desc_sw128(sL + k * 32);
              umma_ss(d_tmem + static_cast<uint32_t>(p.bn), adesc, ldesc, idesc_l, acc);
            }
          }
          umma_commit(bar_empty + 8u * stage);  // frees the smem stage once these MMAs retire
          if (kb == p.kblocks - 1) umma_commit(bar_tfull + 8u * as);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mt = t % p.num_m_tiles;
      const int nt = t / p.num_m_tiles;
      mbar_wait(bar_tfull + 8u * as, aphase);
      tc_fence_after();
      const int m = mt * kBM + q * 32 + lane;
      const bool row_ok = m < p.M;
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * 256u;
      const int n_base = nt * p.ncols_out;
      const __nv_bfloat16* rb =
          (p.flags & SB200_EPI_ROWBIAS)
              ? p.rowbias + static_cast<size_t>(row_ok ? m / p.rows_per_batch : 0) * p.Nout
              : nullptr;
      float tl[8];
      int cur_group = -1;
      const float lscale =
          has_lora ? (p.lora_scale_dev ? p.lora_scale * __ldg(p.lora_scale_dev) : p.lora_scale) : 0.f;
      for (int c = 0; c < p.ncols_out; c += 16) {
        const int n = n_base + c;
        if (n >= p.Nout) break;  // warp-uniform
        uint32_t v[16];
        uint32_t g[16];
        tmem_ld_x16(taddr + c, v);
        if (geglu) tmem_ld_x16(taddr + (p.bn >> 1) + c, g);
        if (has_lora) {
          const int grp = n / p.lora_group_n;
          if (grp != cur_group) {
            cur_group = grp;
            uint32_t tv[8];
            tmem_ld_x8(taddr + p.bn + grp * p.lora_r, tv);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) tl[j] = __uint_as_float(tv[j]) * lscale;
          }
        }
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (has_lora) {
          const __nv_bfloat16* up = p.lora_up + static_cast<size_t>(n) * p.lora_r;
          if (p.lora_r == 4)
            lora_apply<4>(f, tl, up);
          else
            lora_apply<8>(f, tl, up);
        }
        if (p.flags & SB200_EPI_BIAS) {
          const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n);
          const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
          const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[2 * i] += bf16_lo(bw[i]);
            f[2 * i + 1] += bf16_hi(bw[i]);
          }
        }
        if (rb) {
          const uint4* bp = reinterpret_cast<const uint4*>(rb + n);
          const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
          const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[2 * i] += bf16_lo(bw[i]);
            f[2 * i + 1] += bf16_hi(bw[i]);
          }
        }
        if (geglu) {
          float gb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) gb[i] = 0.f;
          if (p.flags & SB200_EPI_BIAS) {
            const uint4* bp = reinterpret_cast<const uint4*>(p.bias + p.Nout + n);
            const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
            const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              gb[2 * i] = bf16_lo(bw[i]);
              gb[2 * i + 1] = bf16_hi(bw[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] *= gelu_erf_f(__uint_as_float(g[i]) + gb[i]);
        }
        if (row_ok) {
          if (p.flags & SB200_EPI_RESID) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.resid + static_cast<size_t>(m) * p.ldr + n);
            const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f[2 * i] += bf16_lo(rw[i]);
              f[2 * i + 1] += bf16_hi(rw[i]);
            }
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]);
          o0.y = pack_bf16x2(f[2], f[3]);
          o0.z = pack_bf16x2(f[4], f[5]);
          o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]);
          o1.y = pack_bf16x2(f[10], f[11]);
          o1.z = pack_bf16x2(f[12], f[13]);
          o1.w = pack_bf16x2(f[14], f[15]);
          uint4* op = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(m) * p.ldo + n);
          op[0] = o0;
          op[1] = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8u * as);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
static int pick_bn(int m_tiles, int n_cols, int max_bn, int step, int num_sms, int kblocks) {
  // minimise waves x (tile time); tile time ~ kblocks * max(bn, 64) cycles*2 + a fixed per-tile cost
  long best_cost = -1;
  int best = step;
  for (int bn = step; bn <= max_bn; bn += step) {
    const int n_tiles = (n_cols + bn - 1) / bn;
    const long tiles = static_cast<long>(m_tiles) * n_tiles;
    const long waves = (tiles + num_sms - 1) / num_sms;
    const long tile_cost = static_cast<long>(kblocks) * (bn > 96 ? bn : 96) + 600;
    const long cost = waves * tile_cost;
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && bn > best)) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

static int launch_gemm(Ctx* ctx, cudaStream_t stream, GemmParams& p) {
  const int lora_bytes = (p.flags & SB200_EPI_LORA) ? p.lora_rt * 128 : 0;
  p.stage_bytes = kBM * 128 + p.bn * 128 + lora_bytes;
  int stages = (kSmemBudget - kBarRegion - 1024) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(SB200_ERR_INVALID, "gemm: tile does not fit shared memory");
  p.stages = stages;
  const int smem = kBarRegion + 1024 + stages * p.stage_bytes;
  if (!ctx->gemm_attr_set) {
    SB200_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kSmemBudget));
    ctx->gemm_attr_set = true;
  }
  const int total = p.num_m_tiles * p.num_n_tiles;
  const int grid = total < ctx->num_sms ? total : ctx->num_sms;
  gemm_kernel<<<grid, kGemmThreads, smem, stream>>>(p);
  SB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int check_lora(const sb200_lora* l, int N, int max_rt) {
  SB200_REQUIRE(l->down && l->up, "lora: NULL weights");
  SB200_REQUIRE(l->r == 4 || l->r == 8, "lora: rank %d unsupported by the fused epilogue (4 or 8)", l->r);
  SB200_REQUIRE(l->rt == 16 || l->rt == 32, "lora: rt must be 16 or 32");
  SB200_REQUIRE(l->group_n > 0 && l->group_n % 16 == 0, "lora: group_n must be a multiple of 16");
  const int groups = (N + l->group_n - 1) / l->group_n;
  SB200_REQUIRE(groups * l->r <= l->rt, "lora: %d groups of rank %d exceed rt=%d", groups, l->r, l->rt);
  (void)max_rt;
  return 0;
}

}  // namespace sb200

using namespace sb200;

extern "C" int sb200_gemm(void* handle, void* stream, const void* x0, int ldx0, const void* x1, int ldx1,
                          int K0, const void* w, int ldw, void* out, int ldo, int M, int N, int K,
                          int flags, const void* bias, const void* rowbias, int rows_per_batch,
                          const void* resid, int ldr, const sb200_lora* lora, int bn) {
  Ctx* ctx = as_ctx(handle);
  SB200_REQUIRE(ctx, "gemm: NULL handle");
  SB200_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad dims M=%d N=%d K=%d", M, N, K);
  SB200_REQUIRE(N % 16 == 0, "gemm: N=%d must be a multiple of 16", N);
  SB200_REQUIRE(K % 8 == 0 && ldx0 % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0,
                "gemm: K and leading dims must be multiples of 8 elements");
  const bool split = x1 != nullptr;
  if (!split) K0 = K;
  SB200_REQUIRE(!split || (K0 % 64 == 0 && K0 > 0 && K0 < K && ldx1 % 8 == 0),
                "gemm: split K0=%d must be a multiple of 64 inside (0, K)", K0);
  const bool geglu = flags & SB200_EPI_GEGLU;
  const bool has_lora = flags & SB200_EPI_LORA;
  SB200_REQUIRE(!(geglu && has_lora), "gemm: GEGLU and LORA cannot be combined");
  SB200_REQUIRE(!geglu || N % 32 == 0, "gemm: GEGLU needs N %% 32 == 0");
  SB200_REQUIRE(!(flags & SB200_EPI_BIAS) || bias, "gemm: BIAS without bias");
  SB200_REQUIRE(!(flags & SB200_EPI_ROWBIAS) || (rowbias && rows_per_batch > 0), "gemm: ROWBIAS args");
  SB200_REQUIRE(!(flags & SB200_EPI_RESID) || (resid && ldr % 8 == 0), "gemm: RESID args");
  SB200_REQUIRE(!has_lora || lora, "gemm: LORA without sb200_lora");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.K = K;
  p.flags = flags;
  p.Nout = geglu ? N / 2 : N;
  p.num_m_tiles = (M + kBM - 1) / kBM;
  p.kblocks = (K + kBK - 1) / kBK;
  p.a_mode = 0;
  p.kb_split = split ? K0 / kBK : p.kblocks;
  p.cb_total = p.kblocks;
  int max_bn = 256;
  if (has_lora) {
    int st = check_lora(lora, N, 32);
    if (st) return st;
    max_bn = 256 - lora->rt;
    p.lora_up = static_cast<const __nv_bfloat16*>(lora->up);
    p.lora_r = lora->r;
    p.lora_rt = lora->rt;
    p.lora_group_n = lora->group_n;
    p.lora_scale = lora->scale;
    p.lora_scale_dev = lora->scale_dev;
  }
  const int step = geglu ? 32 : 16;
  if (bn <= 0) {
    bn = pick_bn(p.num_m_tiles, geglu ? N : N, max_bn, step, ctx->num_sms, p.kblocks);
  }
  SB200_REQUIRE(bn % step == 0 && bn >= step && bn <= max_bn, "gemm: bn=%d invalid (step %d, max %d)", bn,
                step, max_bn);
  p.bn = bn;
  p.ncols_out = geglu ? bn / 2 : bn;
  p.num_n_tiles = (p.Nout + p.ncols_out - 1) / p.ncols_out;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.rowbias = static_cast<const __nv_bfloat16*>(rowbias);
  p.rows_per_batch = rows_per_batch;
  p.resid = static_cast<const __nv_bfloat16*>(resid);
  p.ldr = ldr;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;

  int st;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K0), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldx0) * 2};
    const uint32_t box[2] = {kBK, kBM};
    if ((st = make_tmap_bf16(ctx, &p.tmA[0], x0, 2, dims, strides, box))) return st;
  }
  if (split) {
    const uint64_t dims[2] = {static_cast<uint64_t>(K - K0), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldx1) * 2};
    const uint32_t box[2] = {kBK, kBM};
    if ((st = make_tmap_bf16(ctx, &p.tmA[1], x1, 2, dims, strides, box))) return st;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {kBK, static_cast<uint32_t>(geglu ? bn / 2 : bn)};
    if ((st = make_tmap_bf16(ctx, &p.tmB, w, 2, dims, strides, box))) return st;
  }
  if (has_lora) {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(lora->rt)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {kBK, static_cast<uint32_t>(lora->rt)};
    if ((st = make_tmap_bf16(ctx, &p.tmL, lora->down, 2, dims, strides, box))) return st;
  }
  return launch_gemm(ctx, static_cast<cudaStream_t>(stream), p);
}

extern "C" int sb200_conv3x3(void* handle, void* stream, const void* x0, int ldx0, const void* x1,
                             int ldx1, int C0, int C1, const void* w, void* out, int ldo, int B, int Hin,
                             int Win, int Cout, int stride, int flags, const void* bias,
                             const void* rowbias, const void* resid, int ldr, const sb200_lora* lora,
                             int bn) {
  Ctx* ctx = as_ctx(handle);
  SB200_REQUIRE(ctx, "conv3x3: NULL handle");
  SB200_REQUIRE(stride == 1 || stride == 2, "conv3x3: stride %d", stride);
  SB200_REQUIRE(B > 0 && Hin > 0 && Win > 0 && Hin % stride == 0 && Win % stride == 0, "conv3x3: dims");
  if (!x1) C1 = 0;
  SB200_REQUIRE(C0 > 0 && C0 % 64 == 0 && C1 % 64 == 0, "conv3x3: C0=%d C1=%d must be multiples of 64", C0,
                C1);
  SB200_REQUIRE(Cout % 16 == 0, "conv3x3: Cout=%d must be a multiple of 16", Cout);
  SB200_REQUIRE(!(flags & SB200_EPI_GEGLU), "conv3x3: GEGLU unsupported");
  SB200_REQUIRE(stride == 1 || C1 == 0, "conv3x3: stride 2 takes a single source");
  SB200_REQUIRE(ldx0 % 8 == 0 && ldo % 8 == 0 && (C1 == 0 || ldx1 % 8 == 0), "conv3x3: leading dims");
  const bool has_lora = flags & SB200_EPI_LORA;
  SB200_REQUIRE(!has_lora || lora, "conv3x3: LORA without sb200_lora");
  SB200_REQUIRE(!(flags & SB200_EPI_BIAS) || bias, "conv3x3: BIAS without bias");
  SB200_REQUIRE(!(flags & SB200_EPI_ROWBIAS) || rowbias, "conv3x3: ROWBIAS without rowbias");
  SB200_REQUIRE(!(flags & SB200_EPI_RESID) || (resid && ldr % 8 == 0), "conv3x3: RESID args");
  const int H = Hin / stride, W = Win / stride;
  // 128 output pixels per tile = bb images x bh rows x bw columns, contiguous in NHWC order
  int bw, bh, bb;
  if (W >= 128) {
    SB200_REQUIRE(W % 128 == 0, "conv3x3: W=%d must be a multiple of 128 or divide 128", W);
    bw = 128, bh = 1, bb = 1;
  } else {
    SB200_REQUIRE(128 % W == 0, "conv3x3: W=%d must divide 128", W);
    bw = W;
    const int rows = 128 / W;
    if (rows <= H) {
      SB200_REQUIRE(H % rows == 0, "conv3x3: H=%d must be a multiple of %d", H, rows);
      bh = rows, bb = 1;
    } else {
      SB200_REQUIRE(rows % H == 0, "conv3x3: H=%d must divide %d", H, rows);
      bh = H, bb = rows / H;
    }
  }
  const int Cin = C0 + C1;
  const int M = B * H * W;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = Cout;
  p.K = 9 * Cin;
  p.flags = flags;
  p.Nout = Cout;
  p.num_m_tiles = (M + kBM - 1) / kBM;
  p.cb_total = Cin / kBK;
  p.kblocks = 9 * p.cb_total;
  p.a_mode = stride == 1 ? 1 : 2;
  p.kb_split = C0 / kBK;
  p.H = H;
  p.W = W;
  int max_bn = 256;
  if (has_lora) {
    int st = check_lora(lora, Cout, 32);
    if (st) return st;
    max_bn = 256 - lora->rt;
    p.lora_up = static_cast<const __nv_bfloat16*>(lora->up);
    p.lora_r = lora->r;
    p.lora_rt = lora->rt;
    p.lora_group_n = lora->group_n;
    p.lora_scale = lora->scale;
    p.lora_scale_dev = lora->scale_dev;
  }
  if (bn <= 0) bn = pick_bn(p.num_m_tiles, Cout, max_bn, 16, ctx->num_sms, p.kblocks);
  SB200_REQUIRE(bn % 16 == 0 && bn >= 16 && bn <= max_bn, "conv3x3: bn=%d invalid", bn);
  p.bn = bn;
  p.ncols_out = bn;
  p.num_n_tiles = (Cout + bn - 1) / bn;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.rowbias = static_cast<const __nv_bfloat16*>(rowbias);
  p.rows_per_batch = H * W;
  p.resid = static_cast<const __nv_bfloat16*>(resid);
  p.ldr = ldr;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;

  int st;
  const uint32_t box[4] = {kBK, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh),
                           static_cast<uint32_t>(bb)};
  if (stride == 1) {
    const void* src[2] = {x0, x1};
    const int ld[2] = {ldx0, ldx1};
    const int cs[2] = {C0, C1};
    for (int s = 0; s < (C1 ? 2 : 1); ++s) {
      const uint64_t dims[4] = {static_cast<uint64_t>(cs[s]), static_cast<uint64_t>(Win),
                                static_cast<uint64_t>(Hin), static_cast<uint64_t>(B)};
      const uint64_t pix = static_cast<uint64_t>(ld[s]) * 2;
      const uint64_t strides[3] = {pix, pix * Win, pix * Win * Hin};
      if ((st = make_tmap_bf16(ctx, &p.tmA[s], src[s], 4, dims, strides, box))) return st;
    }
  } else {
    // parity planes: plane (ph, pw) holds input pixels (2*hh + ph, 2*ww + pw)
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const uint64_t pix = static_cast<uint64_t>(ldx0) * 2;
        const uint8_t* basep = static_cast<const uint8_t*>(x0) + (static_cast<uint64_t>(ph) * Win + pw) * pix;
        const uint64_t dims[4] = {static_cast<uint64_t>(C0), static_cast<uint64_t>(W),
                                  static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
        const uint64_t strides[3] = {pix * 2, pix * Win * 2, pix * Win * Hin};
        if ((st = make_tmap_bf16(ctx, &p.tmA[ph * 2 + pw], basep, 4, dims, strides, box))) return st;
      }
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(Cout)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t wbox[2] = {kBK, static_cast<uint32_t>(bn)};
    if ((st = make_tmap_bf16(ctx, &p.tmB, w, 2, dims, strides, wbox))) return st;
  }
  if (has_lora) {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(lora->rt)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t lbox[2] = {kBK, static_cast<uint32_t>(lora->rt)};
    if ((st = make_tmap_bf16(ctx, &p.tmL, lora->down, 2, dims, strides, lbox))) return st;
  }
  return launch_gemm(ctx, static_cast<cudaStream_t>(stream), p);
}
