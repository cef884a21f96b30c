// Multi-head self-attention of the UNFUSED encoder path (training forward / backward, D = 384 inference) on the 5th-generation
// tensor cores: TMA tile loads, tcgen05.mma with accumulators (and the probabilities) in tensor memory, TMA stores.
// T = 64 tokens per image, head dim 64 (docs/VIT_SPEC.md); replaces the round-1 mma.sync kernels.
//
//   forward : S = Q K^T, P = softmax(S / 8), O = P V                         qkv [B*64, 3D] (q | k | v, head-major) -> out [B*64, D]
//   backward: recompute S, P;  dP = dO V^T;  dz = P o (dP - rowsum(dP o P));  dV = P^T dO;  dQ = dz K / 8;  dK = dz^T Q / 8
//
// Work item = (128-token row tile = TWO images, head).  One M = 128 instruction computes both images' 64 x 64 score blocks as the
// diagonal of a 128 x 128 product; the off-diagonal blocks are never read and their probabilities are written as zeros, so the
// products that contract over the key / query index (P V, P^T dO, dz^T Q, dz K) stay per image.  That doubles the tensor work of
// the attention core, which is irrelevant here: the kernels are HBM-bound (forward 64 KB, backward 112 KB per item against
// 512 / 1280 tensor cycles) and are built around keeping TMA loads in flight -- persistent CTAs, multi-stage operand ring.
//
// Warp roles: 0 TMA loader, 1 TMEM allocator + MMA issuer, 2 TMA store warp, 4.. compute warps (softmax / gradient epilogues).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace atc {
constexpr int TM = 128, DH = 64;
constexpr int TILE = TM * DH * 2;                 // one [128 x 64] bf16 operand tile = 16 KB
constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;   // softmax(s / 8) evaluated as 2^((s - max) * log2(e) / 8)

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of the 16-byte chunk holding columns [8*c8, 8*c8+8) of row `row` in a 128B-swizzled [rows x 64] bf16 tile
__device__ __forceinline__ uint32_t swz(int row, int c8) { return (uint32_t)row * 128u + ((((uint32_t)c8) ^ ((uint32_t)row & 7u)) << 4); }
}  // namespace atc

// =====================================================================================================================
// forward
// =====================================================================================================================
namespace atf {
using namespace atc;
constexpr int NST = 3;                             // operand ring stages (Q, K, V tiles: 48 KB each)
constexpr int STAGE_BYTES = 3 * TILE;
constexpr int OFF_RING = 0;
constexpr int OFF_OUT = OFF_RING + NST * STAGE_BYTES;   // 2 output staging tiles
constexpr int OFF_BAR = OFF_OUT + 2 * TILE;
constexpr int OFF_MISC = OFF_BAR + 256;
constexpr int MISC_BYTES = 2 * 128 * 2 * 8;        // softmax (max, sum) partials, double buffered: [2][128][2] float2
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int W_LOAD = 0, W_MMA = 1, W_STORE = 2, W_SM0 = 4, N_SM = 8;
constexpr int THREADS = 32 * (W_SM0 + N_SM);       // 384
constexpr int TMEM_COLS = 512;
constexpr int COL_S = 0, COL_O = 256;              // S[2] @ 0, 128 ; O[2] @ 256, 320
enum { B_FULL = 0, B_EMPTY = B_FULL + NST, B_SFULL = B_EMPTY + NST, B_PREADY = B_SFULL + 2, B_OFULL = B_PREADY + 2, B_OFREE = B_OFULL + 2,
       B_STAGED = B_OFREE + 2, B_STFREE = B_STAGED + 2, B_TMEMSLOT = B_STFREE + 2, B_COUNT };
static_assert(B_COUNT * 8 <= 256, "barrier area");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace atf

struct AttnTcParams {
  int tiles;      // 128-token row tiles (ceil(B / 2))
  int heads;
  int D;
  float* dbqkv;   // backward: [3 D] += column sums of dqkv (nullable)
};

__global__ void __launch_bounds__(atf::THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut, const AttnTcParams p) {
  using namespace atf;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  float2* sm_part = reinterpret_cast<float2*>(sptr + OFF_MISC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.tiles * p.heads;
  const int ni = (int)blockIdx.x < items ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto item_of = [&](int n) { return (int)blockIdx.x + n * (int)gridDim.x; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmOut);
    for (int i = 0; i < NST; ++i) { mbar_init(bar(B_FULL + i), 1); mbar_init(bar(B_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_SFULL + i), 1); mbar_init(bar(B_PREADY + i), N_SM); mbar_init(bar(B_OFULL + i), 1); mbar_init(bar(B_OFREE + i), 4);
      mbar_init(bar(B_STAGED + i), 4); mbar_init(bar(B_STFREE + i), 1);
    }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(bar(B_TMEMSLOT));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMSLOT)));

  if (warp == W_LOAD) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int n = 0; n < ni; ++n) {
        const int w = item_of(n), tile = w / p.heads, h = w % p.heads;
        mbar_wait_guard(bar(B_EMPTY + s), ph ^ 1);
        mbar_arrive_expect_tx(bar(B_FULL + s), STAGE_BYTES);
        const uint32_t dst = sbase + OFF_RING + s * STAGE_BYTES;
        for (int part = 0; part < 3; ++part) tma_load_2d(dst + part * TILE, &tmQKV, part * p.D + h * DH, tile * TM, bar(B_FULL + s));
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    if (ni > 0) {
      constexpr uint32_t id_s = umma_idesc_bf16(TM, 128, false, false);
      constexpr uint32_t id_pv = umma_idesc_bf16(TM, DH, false, true);        // B = V, MN-major (rows = keys)
      int s_s = 0; uint32_t ph_s = 0;      // ring position of the next S product
      int s_pv = 0;                        // ring position of the next P.V product
      auto issue_s = [&](int n) {
        const uint32_t b = n & 1;
        mbar_wait_guard(bar(B_FULL + s_s), ph_s);
        tc_fence_after();
        const uint32_t q = sbase + OFF_RING + s_s * STAGE_BYTES;
        const uint32_t la = umma_desc_lo(q), lb = umma_desc_lo(q + TILE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + COL_S + b * 128, umma_desc_from_lo(la + 2 * k), umma_desc_from_lo(lb + 2 * k), id_s, k ? 1u : 0u);
          umma_commit(bar(B_SFULL + b));
        }
        __syncwarp();
        if (++s_s == NST) { s_s = 0; ph_s ^= 1; }
      };
      issue_s(0);
      for (int n = 0; n < ni; ++n) {
        const uint32_t b = n & 1, use = (n >> 1) & 1;
        if (n + 1 < ni) issue_s(n + 1);              // the next item's scores are computed while this item's softmax runs
        mbar_wait_guard(bar(B_PREADY + b), use);
        mbar_wait_guard(bar(B_OFREE + b), use ^ 1);  // the epilogue of item n-2 has drained O[b]
        tc_fence_after();
        const uint32_t lv = umma_desc_lo(sbase + OFF_RING + s_pv * STAGE_BYTES + 2 * TILE, 8192);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)                 // 16 keys per step: P columns 8k.. (bf16 pairs), V rows 16k.. (2048 B)
            umma_bf16_ts(tmem_base + COL_O + b * DH, tmem_base + COL_S + b * 128 + 8 * k, umma_desc_from_lo(lv + k * 128), id_pv, k ? 1u : 0u);
          umma_commit(bar(B_OFULL + b));
          umma_commit(bar(B_EMPTY + s_pv));           // Q, K, V of this item are dead once P.V has retired
        }
        __syncwarp();
        if (++s_pv == NST) s_pv = 0;
      }
    }
  } else if (warp == W_STORE) {
    if (lane == 0) {
      for (int n = 0; n < ni; ++n) {
        const int w = item_of(n), tile = w / p.heads, h = w % p.heads;
        const uint32_t b = n & 1;
        mbar_wait_guard(bar(B_STAGED + b), (n >> 1) & 1);
        tma_store_2d(&tmOut, sbase + OFF_OUT + b * TILE, h * DH, tile * TM);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(B_STFREE + b));
      }
      bulk_wait0();
    }
  } else if (warp >= W_SM0) {
    // softmax: thread = (query row, half of its image's 64 keys); the half-0 thread of a row also runs its O epilogue
    const int quad = warp & 3, half = (warp - W_SM0) >> 2;
    const int row = quad * 32 + lane, img = row >> 6;
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    for (int n = 0; n < ni; ++n) {
      const uint32_t b = n & 1, use = (n >> 1) & 1;
      mbar_wait_guard(bar(B_SFULL + b), use);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld_32x32(tmem_base + tm_lane + COL_S + b * 128 + img * 64 + half * 32, sv);
      tmem_ld_wait();
      float mx = __uint_as_float(sv[0]);
#pragma unroll
      for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = ex2a((__uint_as_float(sv[i]) - mx) * kScaleLog2);
        sum += e;
        sv[i] = __float_as_uint(e);
      }
      float2* part = sm_part + b * 256;
      part[row * 2 + half] = make_float2(mx, sum);
      asm volatile("bar.sync 2, 256;" ::: "memory");       // also orders every S read of the pair before the in-place P writes
      const float2 other = part[row * 2 + (half ^ 1)];
      const float m = fmaxf(mx, other.x);
      const float f = ex2a((mx - m) * kScaleLog2);
      const float inv = 1.0f / (sum * f + other.y * ex2a((other.x - m) * kScaleLog2));
      uint32_t pw[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pw[i] = pack_bf16(__uint_as_float(sv[2 * i]) * f, __uint_as_float(sv[2 * i + 1]) * f);
      // P (unnormalised, bf16 pairs) over this row's S columns: keys of its own image, zeros for the other image's keys
      tmem_st_32x16(tmem_base + tm_lane + COL_S + b * 128 + img * 32 + half * 16, pw);
      tmem_st_32x16_fill(tmem_base + tm_lane + COL_S + b * 128 + (img ^ 1) * 32 + half * 16, 0u);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PREADY + b));
      if (half == 0) {
        // ---- O epilogue: normalise by 1 / sum, convert, stage the [128 x 64] tile for the TMA store ----
        mbar_wait_guard(bar(B_OFULL + b), use);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_O + b * DH, r0);
        tmem_ld_32x32(tmem_base + tm_lane + COL_O + b * DH + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_OFREE + b));
        mbar_wait_guard(bar(B_STFREE + b), use ^ 1);          // the store of item n-2 has read this staging tile
        uint8_t* st = sptr + OFF_OUT + b * TILE;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(st + swz(row, c)) =
              make_uint4(pack_bf16(__uint_as_float(r0[8 * c]) * inv, __uint_as_float(r0[8 * c + 1]) * inv),
                         pack_bf16(__uint_as_float(r0[8 * c + 2]) * inv, __uint_as_float(r0[8 * c + 3]) * inv),
                         pack_bf16(__uint_as_float(r0[8 * c + 4]) * inv, __uint_as_float(r0[8 * c + 5]) * inv),
                         pack_bf16(__uint_as_float(r0[8 * c + 6]) * inv, __uint_as_float(r0[8 * c + 7]) * inv));
          *reinterpret_cast<uint4*>(st + swz(row, 4 + c)) =
              make_uint4(pack_bf16(__uint_as_float(r1[8 * c]) * inv, __uint_as_float(r1[8 * c + 1]) * inv),
                         pack_bf16(__uint_as_float(r1[8 * c + 2]) * inv, __uint_as_float(r1[8 * c + 3]) * inv),
                         pack_bf16(__uint_as_float(r1[8 * c + 4]) * inv, __uint_as_float(r1[8 * c + 5]) * inv),
                         pack_bf16(__uint_as_float(r1[8 * c + 6]) * inv, __uint_as_float(r1[8 * c + 7]) * inv));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_STAGED + b));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc<atf::TMEM_COLS>(tmem_base);
  }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
namespace atb {
using namespace atc;
constexpr int NST = 2;                             // operand ring stages (Q, K, V, dO tiles: 64 KB each)
constexpr int STAGE_BYTES = 4 * TILE;
constexpr int OFF_RING = 0;
constexpr int OFF_P = OFF_RING + NST * STAGE_BYTES;     // P  [128 i x 128 j] bf16 = two [128 x 64] blocks; later the dV staging tile
constexpr int OFF_DS = OFF_P + 2 * TILE;                // dz [128 i x 128 j] bf16;                       later the dQ | dK staging tiles
constexpr int OFF_BAR = OFF_DS + 2 * TILE;
constexpr int OFF_MISC = OFF_BAR + 256;
constexpr int MAX_HEADS_CS = 8;                         // bias-gradient column sums are kept per CTA for up to 8 heads (ViT-Tiny 3, ViT-S 6)
constexpr int MISC_BYTES = 128 * 2 * 8 + 128 * 2 * 4 + MAX_HEADS_CS * 192 * 4;   // (max, sum) partials, delta partials, column sums [heads][dq|dk|dv][64]
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int W_LOAD = 0, W_MMA = 1, W_STORE = 2, W_C0 = 4, N_C = 12;    // 12 compute warps: softmax-backward on 8 of them, gradient epilogues on all
constexpr int THREADS = 32 * (W_C0 + N_C);         // 512
constexpr int TMEM_COLS = 512;
constexpr int COL_S = 0, COL_DP = 128, COL_DQ = 256, COL_DK = 320, COL_DV = 384;
enum { B_FULL = 0, B_EMPTY = B_FULL + NST, B_SDPFULL = B_EMPTY + NST, B_PDSREADY, B_GFULL, B_GFREE, B_STAGED, B_BUFFREE, B_TMEMSLOT, B_COUNT };
static_assert(B_COUNT * 8 <= 256, "barrier area");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace atb

__global__ void __launch_bounds__(atb::THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQKV,
                   const AttnTcParams p) {
  using namespace atb;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  float2* sm_part = reinterpret_cast<float2*>(sptr + OFF_MISC);
  float* dl_part = reinterpret_cast<float*>(sptr + OFF_MISC + 128 * 2 * 8);
  float* s_cs = dl_part + 128 * 2;                                                  // [heads][3][64]
  const bool want_cs = p.dbqkv != nullptr && p.heads <= MAX_HEADS_CS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.tiles * p.heads;
  const int ni = (int)blockIdx.x < items ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto item_of = [&](int n) { return (int)blockIdx.x + n * (int)gridDim.x; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmDQKV);
    for (int i = 0; i < NST; ++i) { mbar_init(bar(B_FULL + i), 1); mbar_init(bar(B_EMPTY + i), 1); }
    mbar_init(bar(B_SDPFULL), 1); mbar_init(bar(B_PDSREADY), 8); mbar_init(bar(B_GFULL), 1); mbar_init(bar(B_GFREE), N_C);
    mbar_init(bar(B_STAGED), N_C); mbar_init(bar(B_BUFFREE), 1);
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(bar(B_TMEMSLOT));
  for (int i = threadIdx.x; i < MAX_HEADS_CS * 192; i += THREADS) s_cs[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMSLOT)));

  if (warp == W_LOAD) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int n = 0; n < ni; ++n) {
        const int w = item_of(n), tile = w / p.heads, h = w % p.heads;
        mbar_wait_guard(bar(B_EMPTY + s), ph ^ 1);
        mbar_arrive_expect_tx(bar(B_FULL + s), STAGE_BYTES);
        const uint32_t dst = sbase + OFF_RING + s * STAGE_BYTES;
        for (int part = 0; part < 3; ++part) tma_load_2d(dst + part * TILE, &tmQKV, part * p.D + h * DH, tile * TM, bar(B_FULL + s));
        tma_load_2d(dst + 3 * TILE, &tmDO, h * DH, tile * TM, bar(B_FULL + s));
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    if (ni > 0) {
      constexpr uint32_t id_nn = umma_idesc_bf16(TM, 128, false, false);      // S = Q K^T, dP = dO V^T : K-major x K-major, N = 128
      constexpr uint32_t id_tt = umma_idesc_bf16(TM, DH, true, true);         // dV = P^T dO, dK = dz^T Q : MN-major A (rows = queries), MN-major B
      constexpr uint32_t id_nt = umma_idesc_bf16(TM, DH, false, true);        // dQ = dz K : K-major A, MN-major B (rows = keys)
      int s = 0; uint32_t ph = 0;
      for (int n = 0; n < ni; ++n) {
        const uint32_t par = n & 1;
        mbar_wait_guard(bar(B_FULL + s), ph);
        tc_fence_after();
        const uint32_t st = sbase + OFF_RING + s * STAGE_BYTES;
        const uint32_t lq = umma_desc_lo(st), lk = umma_desc_lo(st + TILE), lv = umma_desc_lo(st + 2 * TILE), ldo = umma_desc_lo(st + 3 * TILE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + COL_S, umma_desc_from_lo(lq + 2 * k), umma_desc_from_lo(lk + 2 * k), id_nn, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + COL_DP, umma_desc_from_lo(ldo + 2 * k), umma_desc_from_lo(lv + 2 * k), id_nn, k ? 1u : 0u);
          umma_commit(bar(B_SDPFULL));
        }
        __syncwarp();
        mbar_wait_guard(bar(B_PDSREADY), par);            // P and dz are in shared memory
        mbar_wait_guard(bar(B_GFREE), par ^ 1);           // the gradient accumulators of item n-1 have been drained
        tc_fence_after();
        {
          // operands over the query index i as the contraction (K = 128 rows): 16 rows per step = 2048 B
          const uint32_t lp_mn = umma_desc_lo(sbase + OFF_P, 2 * 8192), lds_mn = umma_desc_lo(sbase + OFF_DS, 2 * 8192);
          const uint32_t lq_mn = umma_desc_lo(st, 8192), ldo_mn = umma_desc_lo(st + 3 * TILE, 8192), lk_mn = umma_desc_lo(st + TILE, 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem_base + COL_DV, umma_desc_from_lo(lp_mn + k * 128), umma_desc_from_lo(ldo_mn + k * 128), id_tt, k ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem_base + COL_DK, umma_desc_from_lo(lds_mn + k * 128), umma_desc_from_lo(lq_mn + k * 128), id_tt, k ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {                  // contraction over the key index j: K-block k >> 2 of the dz tile, 32 B per step
              const uint32_t la = umma_desc_lo(sbase + OFF_DS + (k >> 2) * TILE) + 2 * (k & 3);
              umma_bf16(tmem_base + COL_DQ, umma_desc_from_lo(la), umma_desc_from_lo(lk_mn + k * 128), id_nt, k ? 1u : 0u);
            }
            umma_commit(bar(B_GFULL));
            umma_commit(bar(B_EMPTY + s));
          }
          __syncwarp();
        }
        if (++s == NST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == W_STORE) {
    if (lane == 0) {
      for (int n = 0; n < ni; ++n) {
        const int w = item_of(n), tile = w / p.heads, h = w % p.heads;
        mbar_wait_guard(bar(B_STAGED), n & 1);
        tma_store_2d(&tmDQKV, sbase + OFF_DS, 0 * p.D + h * DH, tile * TM);            // dQ
        tma_store_2d(&tmDQKV, sbase + OFF_DS + TILE, 1 * p.D + h * DH, tile * TM);     // dK
        tma_store_2d(&tmDQKV, sbase + OFF_P, 2 * p.D + h * DH, tile * TM);             // dV
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(B_BUFFREE));                                                    // P / dz buffers may be rewritten
      }
      bulk_wait0();
    }
  } else if (warp >= W_C0) {
    const int cw = warp - W_C0;                      // 0..11
    const int quad = warp & 3, grp = cw >> 2;        // grp: softmax half (0, 1) / gradient (0 dQ, 1 dK, 2 dV)
    const int row = quad * 32 + lane, img = row >> 6;
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    for (int n = 0; n < ni; ++n) {
      const uint32_t par = n & 1;
      if (grp < 2) {
        // ---- softmax backward: thread = (query row i, 32 of its image's 64 keys) ----
        const int half = grp;
        mbar_wait_guard(bar(B_SDPFULL), par);
        tc_fence_after();
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_S + img * 64 + half * 32, sv);
        tmem_ld_32x32(tmem_base + tm_lane + COL_DP + img * 64 + half * 32, dv);
        tmem_ld_wait();
        float mx = __uint_as_float(sv[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = ex2a((__uint_as_float(sv[i]) - mx) * kScaleLog2);
          sum += e;
          sv[i] = __float_as_uint(e);
        }
        sm_part[row * 2 + half] = make_float2(mx, sum);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float2 other = sm_part[row * 2 + (half ^ 1)];
        const float m = fmaxf(mx, other.x);
        const float f = ex2a((mx - m) * kScaleLog2);
        const float pn = f / (sum * f + other.y * ex2a((other.x - m) * kScaleLog2));     // e * pn = P
        float dl = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float pr = __uint_as_float(sv[i]) * pn;
          sv[i] = __float_as_uint(pr);
          dl = fmaf(pr, __uint_as_float(dv[i]), dl);
        }
        dl_part[row * 2 + half] = dl;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        dl += dl_part[row * 2 + (half ^ 1)];
        // the P / dz buffers double as the previous item's output staging tiles
        mbar_wait_guard(bar(B_BUFFREE), par ^ 1);
        uint8_t* pb = sptr + OFF_P + img * TILE;           // block = key index >> 6 = this row's image
        uint8_t* db = sptr + OFF_DS + img * TILE;
        uint8_t* pz = sptr + OFF_P + (img ^ 1) * TILE;     // the other image's keys: zeros
        uint8_t* dz = sptr + OFF_DS + (img ^ 1) * TILE;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pv, gv;
          uint32_t* pw = reinterpret_cast<uint32_t*>(&pv);
          uint32_t* gw = reinterpret_cast<uint32_t*>(&gv);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float p0 = __uint_as_float(sv[8 * c + 2 * i]), p1 = __uint_as_float(sv[8 * c + 2 * i + 1]);
            pw[i] = pack_bf16(p0, p1);
            gw[i] = pack_bf16(p0 * (__uint_as_float(dv[8 * c + 2 * i]) - dl) * 0.125f, p1 * (__uint_as_float(dv[8 * c + 2 * i + 1]) - dl) * 0.125f);
          }
          const uint32_t o = swz(row, half * 4 + c);
          *reinterpret_cast<uint4*>(pb + o) = pv;
          *reinterpret_cast<uint4*>(db + o) = gv;
          *reinterpret_cast<uint4*>(pz + o) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(dz + o) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PDSREADY));
      }
      // ---- gradient epilogue: thread = (token row, one of dQ / dK / dV): 64 accumulator columns -> bf16 -> staging tile ----
      mbar_wait_guard(bar(B_GFULL), par);
      tc_fence_after();
      const uint32_t col = grp == 0 ? COL_DQ : (grp == 1 ? COL_DK : COL_DV);
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem_base + tm_lane + col, r0);
      tmem_ld_32x32(tmem_base + tm_lane + col + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_GFREE));
      // (GFULL implies the products that read P / dz have retired, so the buffers may be overwritten by the staged outputs)
      uint8_t* st = sptr + (grp == 0 ? OFF_DS : (grp == 1 ? OFF_DS + TILE : OFF_P));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        *reinterpret_cast<uint4*>(st + swz(row, c)) =
            make_uint4(pack_bf16(__uint_as_float(r0[8 * c]), __uint_as_float(r0[8 * c + 1])), pack_bf16(__uint_as_float(r0[8 * c + 2]), __uint_as_float(r0[8 * c + 3])),
                       pack_bf16(__uint_as_float(r0[8 * c + 4]), __uint_as_float(r0[8 * c + 5])), pack_bf16(__uint_as_float(r0[8 * c + 6]), __uint_as_float(r0[8 * c + 7])));
        *reinterpret_cast<uint4*>(st + swz(row, 4 + c)) =
            make_uint4(pack_bf16(__uint_as_float(r1[8 * c]), __uint_as_float(r1[8 * c + 1])), pack_bf16(__uint_as_float(r1[8 * c + 2]), __uint_as_float(r1[8 * c + 3])),
                       pack_bf16(__uint_as_float(r1[8 * c + 4]), __uint_as_float(r1[8 * c + 5])), pack_bf16(__uint_as_float(r1[8 * c + 6]), __uint_as_float(r1[8 * c + 7])));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_STAGED));
      if (want_cs) {
        // QKV bias gradient: column sums over the warp's 32 token rows (transposing butterfly), per CTA in shared memory
        const int h = item_of(n) % p.heads;
        float cs[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) cs[i] = __uint_as_float(r0[i]);
        atomicAdd(&s_cs[(h * 3 + grp) * 64 + lane], warp_colsum32(cs, lane));
#pragma unroll
        for (int i = 0; i < 32; ++i) cs[i] = __uint_as_float(r1[i]);
        atomicAdd(&s_cs[(h * 3 + grp) * 64 + 32 + lane], warp_colsum32(cs, lane));
      }
    }
    if (want_cs) {
      asm volatile("bar.sync 3, 384;" ::: "memory");         // the 12 compute warps
      for (int i = threadIdx.x - W_C0 * 32; i < p.heads * 192; i += N_C * 32) {
        const int h = i / 192, part = (i % 192) / 64, c = i % 64;
        atomicAdd(p.dbqkv + part * p.D + h * 64 + c, s_cs[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc<atb::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
static int attn_check(int B, int heads) {
  if (B < 0 || heads < 1 || heads > 64) { set_last_error("attention: bad shape"); return VITMARL_EINVAL; }
  return VITMARL_OK;
}

// qkv [B*64, 3*D] (q | k | v, head-major) -> out [B*64, D]
int launch_attention(cudaStream_t s, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int heads) {
  using namespace atf;
  int rc = attn_check(B, heads);
  if (rc || B == 0) return rc;
  const int D = heads * DH, M = B * 64;
  CUtensorMap tmQKV, tmOut;
  if ((rc = make_tmap_2d_bf16(&tmQKV, qkv, M, 3 * D, (uint64_t)3 * D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmOut, out, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  AttnTcParams p{(M + TM - 1) / TM, heads, D, nullptr};
  cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e);
  const int grid = min(p.tiles * heads, num_sms());
  attn_tc_fwd_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tmQKV, tmOut, p);
  return check_cuda(cudaGetLastError());
}

// qkv, dout [B*64, D] -> dqkv [B*64, 3*D]
int launch_attention_bwd(cudaStream_t s, const __nv_bfloat16* qkv, const __nv_bfloat16* dout, __nv_bfloat16* dqkv, int B, int heads, float* dbqkv) {
  using namespace atb;
  int rc = attn_check(B, heads);
  if (rc || B == 0) return rc;
  const int D = heads * DH, M = B * 64;
  CUtensorMap tmQKV, tmDO, tmDQKV;
  if ((rc = make_tmap_2d_bf16(&tmQKV, qkv, M, 3 * D, (uint64_t)3 * D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmDO, dout, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmDQKV, dqkv, M, 3 * D, (uint64_t)3 * D * 2, TM, 64))) return rc;
  AttnTcParams p{(M + TM - 1) / TM, heads, D, heads <= MAX_HEADS_CS ? dbqkv : nullptr};
  cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e);
  const int grid = min(p.tiles * heads, num_sms());
  attn_tc_bwd_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(tmQKV, tmDO, tmDQKV, p);
  if ((rc = check_cuda(cudaGetLastError()))) return rc;
  if (dbqkv && heads > MAX_HEADS_CS) return launch_colsum(s, dqkv, dbqkv, M, 3 * D);     // beyond the per-CTA table: a separate pass
  return VITMARL_OK;
}

}  // namespace vitmarl
