// Flash-attention forward for sm_100a:  O = softmax(Q K^T * scale) V  per (batch, head), head dim d <= 192
// (SDXL: 64; SD1.x: 40 / 80 / 160).  The head dim is padded to ND x 64 columns by TMA's out-of-bounds zero
// fill: Q/K/V are addressed through 4-D descriptors (d, heads, tokens, batch), so a 64-wide box that sticks out
// of a 40- or 80-wide head reads zeros, never the neighbouring head.
//
// One CTA per (128-query tile, head, batch).  ND = 1: two CTAs resident per SM (256 TMEM columns, ~81 KB smem
// each) so one CTA's tensor-core work overlaps the other's softmax; ND = 2, 3: one CTA per SM.
//   warp 0      TMA producer: Q tile once, then K / V tiles (128 keys x ND x 64) through a ring of stages
//   warp 1      MMA issuer  : S = Q K^T   (tcgen05.mma SS, 128x128x16 per k-step, fp32 in TMEM columns [0,128))
//                             O += P V    (tcgen05.mma TS: P read from TMEM, V MN-major from smem, N = 64 per
//                                          sub-tile of the head dim)
//   warps 2..9  softmax     : one query row per PAIR of threads (warp w and w + 4 split the 128 key columns). Two passes over S in TMEM (row max, then exp2 + bf16
//                             pack), P written back to TMEM columns [128,192) as the A operand of the PV MMA; the
//                             O accumulator (columns [192, 192 + 64 ND)) stays in TMEM and is rescaled lazily
//                             (only when the running max grows by > 2^8).
// Keys beyond Skv (cross-attention: 77) are masked to -inf; TMA zero-fills the out-of-range rows.
//
// Replaces the attention processor that diffusers' Attention calls (xformers memory_efficient_attention /
// torch SDPA; reference: trainscripts/textsliders/train_lora_xl.py:79-80, train_lora.py:68).
#include "common.h"
#include "ptx.cuh"

namespace sb200 {

constexpr int kAttnThreads = 320;  // TMA warp, MMA warp, 8 softmax warps
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;

template <int ND>
struct AttnCfg {
  static constexpr int kStages = ND == 3 ? 1 : 2;
  static constexpr int kSmem = kTileBytes * ND * (1 + 2 * kStages) + 1024 /*barriers*/ + 2048 /*pair exchange*/ + 1024 /*align*/;
  static constexpr int kTmemCols = ND == 1 ? 256 : 512;
  static constexpr int kMinBlocks = ND == 1 ? 2 : 1;
};

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* o;
  float* lse;      // optional [B, heads, Sq]: log2-domain log-sum-exp of scale*log2(e)*S (for attention_bwd.cu)
  int ldo;
  int Sq, Skv;
  int d;           // head dim (multiple of 8)
  int n_kv_tiles;
  int split_issue;   // 1: Q K^T issued by the TMA warp, P V by the MMA warp (default); 0: both by the MMA warp
  float scale_log2;  // scale * log2(e)
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int ND>
__global__ void __launch_bounds__(kAttnThreads, AttnCfg<ND>::kMinBlocks)
    attention_kernel(const __grid_constant__ AttnParams p) {
  constexpr int kStages = AttnCfg<ND>::kStages;
  constexpr uint32_t kStageBytes = kTileBytes * ND;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;

  const uint32_t sQ = base;
  const uint32_t sK = base + kStageBytes;
  const uint32_t sV = sK + kStages * kStageBytes;
  const uint32_t bars = sV + kStages * kStageBytes;
  const uint32_t bar_q = bars;
  const uint32_t bar_kfull = bars + 8;                   // kStages
  const uint32_t bar_kempty = bar_kfull + 8 * kStages;   // kStages
  const uint32_t bar_vfull = bar_kempty + 8 * kStages;
  const uint32_t bar_vempty = bar_vfull + 8 * kStages;
  const uint32_t bar_sfull = bar_vempty + 8 * kStages;
  const uint32_t bar_sfree = bar_sfull + 8;
  const uint32_t bar_pfull = bar_sfree + 8;
  const uint32_t bar_pvdone = bar_pfull + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bars - base) + 512);
  float* xch = reinterpret_cast<float*>(smem + (bars - base) + 1024);  // [2 parities][2 halves][128 rows]

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_kfull + 8 * i, 1);
      mbar_init(bar_kempty + 8 * i, 1);
      mbar_init(bar_vfull + 8 * i, 1);
      mbar_init(bar_vempty + 8 * i, 1);
    }
    mbar_init(bar_sfull, 1);
    mbar_init(bar_sfree, 8);
    mbar_init(bar_pfull, 8);
    mbar_init(bar_pvdone, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), AttnCfg<ND>::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int n = p.n_kv_tiles;
  const int dsteps = (p.d + 15) >> 4;  // 16-wide k-steps of the QK^T contraction that hold data

  // S_t = Q K_t^T.  Tile j + 1 is issued BEFORE P_j V_j, as soon as the softmax warps have read S_j, so the next S is
  // ready when they finish tile j.  One thread can issue a tcgen05.mma only every ~176 cycles whatever its size
  // (tools/micro/umma_bench.cu), and a tile needs 4 + 8 of them, so the two products are issued by DIFFERENT warps
  // (split_issue): warp 0 issues Q K^T right behind its TMA loads, warp 1 issues P V.  (Measured neutral at d = 64:
  // 646 vs 643 TFLOP/s — with two CTAs per SM the SM already has two issuing threads; kept because it shortens the
  // MMA warp's critical path for ND > 1.)
  const uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0);
  auto issue_qk = [&](int t) {
    const int s = t % kStages;
    const uint32_t ph = (t / kStages) & 1;
    mbar_wait(bar_kfull + 8 * s, ph);
    mbar_wait(bar_sfree, (t & 1) ^ 1u);  // softmax finished reading S_{t-1}
    tc_fence_after();
    if (elect_one()) {
      for (int k = 0; k < dsteps; ++k) {
        const uint32_t off = static_cast<uint32_t>(k >> 2) * kTileBytes + static_cast<uint32_t>(k & 3) * 32;
        umma_ss(tmem_base + kColS, umma_desc_sw128(sQ + off), umma_desc_sw128(sK + s * kStageBytes + off), idesc_qk,
                k != 0);
      }
      umma_commit(bar_kempty + 8 * s);
      umma_commit(bar_sfull);
    }
    __syncwarp();
  };

  if (warp == 0) {
    // converged warp, one ELECTed lane issues (a `lane == 0` branch makes ptxas wrap every uniform-datapath
    // instruction in an ELECT / BRA.U.ANY retry loop)
    if (elect_one()) {
      mbar_expect_tx(bar_q, kStageBytes);
#pragma unroll
      for (int i = 0; i < ND; ++i) tma_load_4d(sQ + i * kTileBytes, &p.tmQ, bar_q, i * 64, head, qt * 128, b);
    }
    __syncwarp();
    for (int j = 0; j < n; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (j / kStages) & 1;
      mbar_wait(bar_kempty + 8 * s, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(bar_kfull + 8 * s, kStageBytes);
#pragma unroll
        for (int i = 0; i < ND; ++i)
          tma_load_4d(sK + s * kStageBytes + i * kTileBytes, &p.tmK, bar_kfull + 8 * s, i * 64, head, j * 128, b);
      }
      __syncwarp();
      mbar_wait(bar_vempty + 8 * s, ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(bar_vfull + 8 * s, kStageBytes);
#pragma unroll
        for (int i = 0; i < ND; ++i)
          tma_load_4d(sV + s * kStageBytes + i * kTileBytes, &p.tmV, bar_vfull + 8 * s, i * 64, head, j * 128, b);
      }
      __syncwarp();
      if (p.split_issue) {
        if (j == 0) mbar_wait(bar_q, 0);
        issue_qk(j);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 1);  // B (= V) is MN-major
    mbar_wait(bar_q, 0);
    if (!p.split_issue) issue_qk(0);
    for (int j = 0; j < n; ++j) {
      const int s = j % kStages;
      const uint32_t ph = (j / kStages) & 1;
      if (!p.split_issue && j + 1 < n) issue_qk(j + 1);
      // ---- O += P_j V_j
      mbar_wait(bar_vfull + 8 * s, ph);
      mbar_wait(bar_pfull, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const int kv_left = p.Skv - j * 128;
        const int ksteps = kv_left >= 128 ? 8 : (kv_left + 15) >> 4;  // keys beyond Skv contribute nothing
        for (int k = 0; k < ksteps; ++k) {
#pragma unroll
          for (int i = 0; i < ND; ++i)
            umma_ts(tmem_base + kColO + i * 64, tmem_base + kColP + k * 8,
                    umma_desc_sw128(sV + s * kStageBytes + i * kTileBytes + k * 2048), idesc_pv, (j | k) != 0);
        }
        umma_commit(bar_vempty + 8 * s);
        umma_commit(bar_pvdone);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax / correction / store
    // Two warps per 32-row TMEM lane quarter (warps w and w + 4): each owns 64 of the tile's 128 key columns.  The
    // pair exchanges its row maxima through shared memory once per tile.  (Measured: 551 -> 565 TFLOP/s at S = 4096;
    // a variant that kept the 64 S values in registers to read TMEM once was 40 % slower at 96 registers / thread.)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tl = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t colS = kColS + half * 64, colP = kColP + half * 32;
    const uint32_t pair_bar = 1 + q;
    float m_run = -INFINITY;  // running max, already multiplied by scale*log2e
    float l_run = 0.f;        // this warp's share of the row sum
    for (int j = 0; j < n; ++j) {
      mbar_wait(bar_sfull, j & 1);
      tc_fence_after();
      const int kv_left = p.Skv - j * 128 - half * 64;  // valid keys among this warp's 64 columns (may be <= 0)
      // pass 1: row max over my columns
      float mx = -INFINITY;
      if (kv_left > 0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_x32(tl + colS + c * 32, v);
          tmem_ld_wait();
          if (kv_left >= (c + 1) * 32) {
            // four independent chains: a single running max is a 32-deep dependent chain (4 clk each)
            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]),
                  m3 = __uint_as_float(v[3]);
#pragma unroll
            for (int i = 4; i < 32; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[i]));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
            }
            mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < kv_left) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
      }
      float* xs = xch + ((j & 1) * 2) * 128;
      xs[half * 128 + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mx = fmaxf(mx, xs[(half ^ 1) * 128 + row]);
      mx *= p.scale_log2;
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_run = mx;
      } else if (mx - m_run > 8.f) {
        need = true;
        alpha = fast_exp2(m_run - mx);
        m_run = mx;
      }
      if (j > 0) {
        // P_{j-1} has been consumed and O holds the sum over tiles < j
        mbar_wait(bar_pvdone, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {  // identical in both warps of the pair (same rows, same maxima)
          l_run *= alpha;
#pragma unroll
          for (int c = 0; c < ND; ++c) {
            uint32_t o[32];
            tmem_ld_x32(tl + kColO + c * 64 + half * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x16(tl + kColO + c * 64 + half * 32, o);
            tmem_st_x16(tl + kColO + c * 64 + half * 32 + 16, o + 16);
          }
          tmem_st_wait();
        }
      }
      // pass 2: p = exp2(s*scale*log2e - m), pack to bf16, store as the A operand of the PV MMA
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        if (kv_left > c * 32) {
          tmem_ld_x32(tl + colS + c * 32, v);
          tmem_ld_wait();
        }
        if (c == 1) {
          // every S column of this warp is now in registers: release S for the next QK MMA
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_sfree);
        }
        if (kv_left <= c * 32) continue;  // columns beyond Skv: the PV MMA never reads them (ksteps limit)
        uint32_t pk[16];
        if (kv_left >= (c + 1) * 32) {  // warp-uniform: no masking code on full chunks
#pragma unroll
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // independent partial sums (no 32-deep add chain)
          for (int i = 0; i < 16; i += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m_run));
            const float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m_run));
            const float p2 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 2]), p.scale_log2, -m_run));
            const float p3 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 3]), p.scale_log2, -m_run));
            s0 += p0;
            s1 += p1;
            s2 += p2;
            s3 += p3;
            pk[i] = pack_bf16x2(p0, p1);
            pk[i + 1] = pack_bf16x2(p2, p3);
          }
          lsum += (s0 + s1) + (s2 + s3);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m_run));
            float p1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m_run));
            if (c * 32 + 2 * i >= kv_left) p0 = 0.f;
            if (c * 32 + 2 * i + 1 >= kv_left) p1 = 0.f;
            lsum += p0 + p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
        }
        tmem_st_x16(tl + colP + c * 16, pk);
      }
      l_run += lsum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pfull);
    }
    // ---- finalize: O / l -> bf16 -> global (only the first d columns of the padded head); each warp of the pair
    // writes its 32 of every 64 O columns
    {
      float* xs = xch + ((n & 1) * 2) * 128;
      xs[half * 128 + row] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      l_run += xs[(half ^ 1) * 128 + row];
    }
    mbar_wait(bar_pvdone, (n - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int srow = qt * 128 + row;
    if (half == 0 && p.lse != nullptr && srow < p.Sq)
      p.lse[(static_cast<size_t>(b) * gridDim.y + head) * p.Sq + srow] = m_run + log2f(l_run);
    __nv_bfloat16* op = p.o + (static_cast<size_t>(b) * p.Sq + srow) * p.ldo + head * p.d;
#pragma unroll
    for (int c = 0; c < ND; ++c) {
      const int col0 = c * 64 + half * 32;
      if (col0 < p.d) {  // warp-uniform
        uint32_t o[32];
        tmem_ld_x32(tl + kColO + col0, o);
        tmem_ld_wait();
        if (srow < p.Sq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (col0 + i * 8 < p.d) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
              w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
              w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
              w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
              *reinterpret_cast<uint4*>(op + col0 + i * 8) = w;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AttnCfg<ND>::kTmemCols);
  }
}

template <int ND>
static int launch_attention(Ctx* ctx, cudaStream_t stream, const AttnParams& p, dim3 grid) {
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[ND]) {
    SB200_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          AttnCfg<ND>::kSmem));
    attr_set[ND] = true;
  }
  SB200_CUDA_CHECK(launch_pdl(attention_kernel<ND>, grid, dim3(kAttnThreads), AttnCfg<ND>::kSmem, stream, p));
  SB200_CUDA_CHECK(cudaGetLastError());
  (void)ctx;
  return 0;
}

}  // namespace sb200

using namespace sb200;

extern "C" int sb200_attention(void* handle, void* stream, const void* q, int ldq, const void* k, int ldk,
                               const void* v, int ldv, void* o, int ldo, int B, int heads, int Sq, int Skv,
                               int head_dim, float scale, float* lse) {
  Ctx* ctx = as_ctx(handle);
  SB200_REQUIRE(ctx, "attention: NULL handle");
  SB200_REQUIRE(B > 0 && heads > 0 && Sq > 0 && Skv > 0, "attention: bad dims");
  SB200_REQUIRE(head_dim >= 8 && head_dim <= 192 && head_dim % 8 == 0,
                "attention: head dim %d unsupported (multiple of 8, <= 192)", head_dim);
  SB200_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
                "attention: leading dims must be multiples of 8");
  SB200_REQUIRE(ldq >= heads * head_dim && ldk >= heads * head_dim && ldv >= heads * head_dim &&
                    ldo >= heads * head_dim,
                "attention: heads=%d x head_dim=%d exceeds a row", heads, head_dim);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  int st;
  const uint32_t box[4] = {64, 1, 128, 1};
  const uint64_t hs = static_cast<uint64_t>(head_dim) * 2;  // byte stride between heads
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(head_dim), static_cast<uint64_t>(heads),
                              static_cast<uint64_t>(Sq), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {hs, static_cast<uint64_t>(ldq) * 2, static_cast<uint64_t>(ldq) * 2 * Sq};
    if ((st = make_tmap_bf16(ctx, &p.tmQ, q, 4, dims, strides, box))) return st;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(head_dim), static_cast<uint64_t>(heads),
                              static_cast<uint64_t>(Skv), static_cast<uint64_t>(B)};
    const uint64_t sk[3] = {hs, static_cast<uint64_t>(ldk) * 2, static_cast<uint64_t>(ldk) * 2 * Skv};
    const uint64_t sv[3] = {hs, static_cast<uint64_t>(ldv) * 2, static_cast<uint64_t>(ldv) * 2 * Skv};
    if ((st = make_tmap_bf16(ctx, &p.tmK, k, 4, dims, sk, box))) return st;
    if ((st = make_tmap_bf16(ctx, &p.tmV, v, 4, dims, sv, box))) return st;
  }
  p.o = static_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.ldo = ldo;
  p.Sq = Sq;
  p.Skv = Skv;
  p.d = head_dim;
  p.n_kv_tiles = (Skv + 127) / 128;
  {
    static const bool split = [] {
      const char* e = getenv("SB200_ATTN_SPLIT");  // "0" = single issuing warp (same-box A/B)
      return !(e && e[0] == '0');
    }();
    p.split_issue = split ? 1 : 0;
  }
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Sq + 127) / 128, heads, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pdl_hint() = static_cast<long long>(grid.x) * grid.y * grid.z <= 4LL * ctx->num_sms;
  if (head_dim <= 64) return launch_attention<1>(ctx, s, p, grid);
  if (head_dim <= 128) return launch_attention<2>(ctx, s, p, grid);
  return launch_attention<3>(ctx, s, p, grid);
}
