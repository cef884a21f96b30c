// Fused transformer MLP block for D = 192 (ViT-Tiny), inference path, CTA-PAIR version (tcgen05 cta_group::2):
//     out = x + FC2( gelu_tanh( FC1( LayerNorm(x) ) ) )
// A thread-block cluster of two CTAs (one TPC) owns a 256-token tile: each CTA stages its own 128 token rows (the A
// operands) and HALF of every weight K-block (the B operands); the leader CTA issues M = 256 `tcgen05.mma` instructions
// that read A from the CTA owning the rows and B from both CTAs.  Compared with the 1-CTA kernel (fused_mlp.cu):
//   * every weight byte crosses L2 -> SM once per 256 tokens (half the traffic) and the per-instruction smem operand
//     read drops from 8-10 KB to 6-7 KB per SM, so the MMAs run at the tensor-pipe rate instead of the smem rate;
//   * the weight rings shrink by half, which pays for a SECOND x-tile buffer: the next tile's TMA load and LayerNorm
//     run while the tensor pipe is still busy with the current tile (no prologue / epilogue bubbles between tiles);
//   * the CUDA-core work per token is cut: LayerNorm scale/shift are folded into W1 / b1 on the fly
//     (vit_fold.cu: W1' = W1 . diag(gamma), b1' = b1 + W1 . beta), statistics use mixed-precision bf16->fp32 FMAs
//     (FHFMA/FHADD, no unpacking), the bias add and GELU run on packed bf16x2 with the 0.5 folded into W2.
//
// Per-CTA warp roles (640 threads): warp 0 x-tile TMA loads + output TMA stores, warp 1 TMEM allocator + (leader only)
// MMA issuer, warp 2 / 3 W1 / W2 weight-ring producers, warps 4-11 GELU epilogue (MUFU-bound), warps 12-19 LayerNorm of the
// next tile + final epilogue of the previous one (running concurrently with the GELU warps).
// Hidden dimension in 6 chunks of 128: acc1[g&1] = LN(x) . W1'[c]^T (TMEM, double buffered); the GELU epilogue writes
// H = gelu(acc1 + b1') as packed bf16 back into the SAME TMEM columns it was read from (tcgen05.st; thread (row, 32 cols)
// overwrites the first 16 of its own 32 columns = two K=16 steps), and acc2 += H . (W2/2)[:,c]^T runs with the A operand
// IN TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): the hidden activation never touches shared memory, which is the
// bandwidth that bounds D = 192 blocks (an SS-mode MMA re-reads its 4 KB A tile from smem for every N <= 192 step).
// The issue order FC1(g+1), FC2(g) runs over a GLOBAL chunk index g = tile * 6 + c, so the pipeline never drains at
// tile boundaries; tcgen05.mma instructions execute in issue order, so FC1(g+2) overwriting acc1[g&1] after FC2(g) has
// read H from it needs no barrier.
//
// Barrier protocol (leader = cluster rank 0; all barriers exist at the same offset in both CTAs):
//   local      : XFULL[2] (x tile landed), OUTREADY (output tile staged in smem)
//   multicast  : W1EMPTY[], W2EMPTY[], ACC1FULL[2], ACC2FULL  (tcgen05.commit of the leader arrives on BOTH CTAs' barrier)
//   on leader  : W1FULL[], W2FULL[] (TMA bytes of both CTAs), XNREADY, HREADY[2], ACC2EMPTY
//                (one arrival per compute warp of BOTH CTAs, remote `mbarrier.arrive.release.cluster`)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace fmlp2 {
constexpr int D = 192, HID = 768, HC = 128, NCHUNK = HID / HC;
constexpr int TM = 128;                                           // token rows per CTA (256 per pair)
constexpr int KB_X = D / 64, KB_H = HC / 64;
constexpr int NB = 2;                                             // acc1 buffers (H aliases them)
constexpr int NXB = 2;                                            // x tile buffers
constexpr int NS1 = 3;                                            // W1 ring, one stage per chunk: this CTA's 64 rows of 3 [128 x 64] K-blocks (24 KB)
constexpr int NS2 = 2;                                            // W2 ring, one stage per chunk: this CTA's 96 rows of 2 [192 x 64] K-blocks (24 KB)
constexpr int S1K = (HC / 2) * 128, S2K = (D / 2) * 128;             // one K-block of a stage (8 KB / 12 KB)
constexpr int S1_BYTES = KB_X * S1K, S2_BYTES = KB_H * S2K;
constexpr int KBLK = TM * 128;                                    // one [128 x 64] bf16 K-block of an A operand = 16 KB
constexpr int X_BYTES = KB_X * KBLK;                              // 48 KB
constexpr int OFF_X = 0;
constexpr int OFF_W1 = OFF_X + NXB * X_BYTES;
constexpr int OFF_W2 = OFF_W1 + NS1 * S1_BYTES;
constexpr int OFF_BAR = OFF_W2 + NS2 * S2_BYTES;
constexpr int OFF_MISC = OFF_BAR + 512;
constexpr int MISC_BYTES = 128 * 4 * 8 + HID * 2 + D * 4;         // LN partials [128][2] float2 (+ spare), b1' (bf16), b2 (packed bf16)
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int NGW = 8, NAW = 8;                                     // GELU warps / aux warps (LayerNorm + final epilogue) per CTA
constexpr int NCW = NGW + NAW;
constexpr int FIRST_CW = 4, FIRST_AUX = FIRST_CW + NGW;
constexpr int THREADS = 32 * (FIRST_CW + NCW);
constexpr int TMEM_COLS = 512;                                    // acc1[2] @ 0,128 ; acc2 @ 256 (192 cols)
constexpr int ACC2_COL = 256;
enum { B_XFULL = 0, B_OUTREADY = B_XFULL + NXB, B_XNREADY, B_ACC2FULL = B_XNREADY + NXB, B_ACC2EMPTY,
       B_ACC1FULL, B_HREADY = B_ACC1FULL + NB,
       B_W1FULL = B_HREADY + NB, B_W1EMPTY = B_W1FULL + NS1, B_W2FULL = B_W1EMPTY + NS1, B_W2EMPTY = B_W2FULL + NS2,
       B_TMEMSLOT = B_W2EMPTY + NS2, B_COUNT };
static_assert(B_COUNT * 8 <= 512, "barrier area");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert((S1K % 1024) == 0 && (S2K % 1024) == 0, "swizzle atoms");
}  // namespace fmlp2

struct FusedMlp2Params {
  int M;                        // tokens
  const __nv_bfloat16* x;       // [M, D] residual stream in (re-read from L2 for the residual add)
  const uint32_t* b1p;          // [HID/2] folded FC1 bias, packed bf16x2
  const float* b2;              // [D]
  float eps;
  long long* dbg;               // optional clock64 timeline of cluster 0 / leader (second tile); null in production
  int inplace;                  // out aliases x: the residual add is done by a TMA reduce-add store (x += delta), no residual load
};

#define FM2_WAIT(acc, call) do { if (p.dbg) { const long long t__ = clock64(); call; if (j == 1) acc += clock64() - t__; } else { call; } } while (0)
#define FM2_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && j == 1 && (threadIdx.x & 31) == 0) p.dbg[(slot)] = clock64(); } while (0)

// x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) on a packed pair = 2 gelu_tanh(x); the 1/2 lives in the FC2 weights
__device__ __forceinline__ uint32_t gelu2_bf16x2(uint32_t x) {
  const uint32_t C0 = 0x3F4C3F4Cu;   // bf16(0.7978845608) x2
  const uint32_t C1 = 0x3D123D12u;   // bf16(0.7978845608 * 0.044715) x2
  const uint32_t x2 = bf16x2_mul(x, x);
  const uint32_t pp = bf16x2_fma(x2, C1, C0);
  const uint32_t u = bf16x2_mul(pp, x);
  uint32_t t;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(u));
  return bf16x2_fma(x, t, x);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(fmlp2::THREADS, 1)
fused_mlp2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const FusedMlp2Params p) {
  using namespace fmlp2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  float2* ln_part = reinterpret_cast<float2*>(sptr + OFF_MISC);                    // [128][4]
  uint32_t* s_b1p = reinterpret_cast<uint32_t*>(sptr + OFF_MISC + 128 * 4 * 8);    // [HID/2] packed bf16x2
  float* s_b2 = reinterpret_cast<float*>(s_b1p + HID / 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_clusters = gridDim.x >> 1, cluster = blockIdx.x >> 1;
  const int pair_tiles = (p.M + 2 * TM - 1) / (2 * TM);
  const int nt = cluster < pair_tiles ? (pair_tiles - cluster + num_clusters - 1) / num_clusters : 0;   // tiles of this cluster
  auto tile_row = [&](int j) { return (cluster + j * num_clusters) * 2 * TM + (int)rank * TM; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmOut);
    for (int i = 0; i < NXB; ++i) mbar_init(bar(B_XFULL + i), 1);
    mbar_init(bar(B_OUTREADY), NAW);
    for (int i = 0; i < NXB; ++i) mbar_init(bar(B_XNREADY + i), 2 * NAW);   // one per x buffer: a LayerNorm may run two tiles ahead of the issuer
    mbar_init(bar(B_ACC2FULL), 1); mbar_init(bar(B_ACC2EMPTY), 2 * NAW);
    for (int i = 0; i < NB; ++i) {
      mbar_init(bar(B_ACC1FULL + i), 1); mbar_init(bar(B_HREADY + i), 2 * NGW);
    }
    for (int i = 0; i < NS1; ++i) { mbar_init(bar(B_W1FULL + i), 1); mbar_init(bar(B_W1EMPTY + i), 1); }
    for (int i = 0; i < NS2; ++i) { mbar_init(bar(B_W2FULL + i), 1); mbar_init(bar(B_W2EMPTY + i), 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar(B_TMEMSLOT)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < HID / 2; i += THREADS) s_b1p[i] = p.b1p[i];
  for (int i = threadIdx.x; i < D / 2; i += THREADS) s_b2[i] = __uint_as_float(pack_bf16(p.b2[2 * i], p.b2[2 * i + 1]));   // packed bf16x2
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMSLOT)));
  // PDL: everything above (and the weight rings of warps 2 / 3, which only read parameters) overlaps the tail of the
  // previous kernel; the residual stream is touched only after it has completed.
  if (warp != 2 && warp != 3) { griddep_wait(); griddep_launch_dependents(); }

  if (warp == 0) {
    // =============================== x tiles in, output tiles out ===============================
    if (lane == 0 && nt > 0) {
      auto load_x = [&](int j) {
        const uint32_t dst = sbase + OFF_X + (j & 1) * X_BYTES, fb = bar(B_XFULL + (j & 1));
        mbar_arrive_expect_tx(fb, X_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(dst + kb * KBLK, &tmX, kb * 64, tile_row(j), fb);
      };
      load_x(0);
      if (nt > 1) load_x(1);
      for (int j = 0; j < nt; ++j) {
        // the final epilogue has staged the output tile in the (dead) x buffer: store it, then reuse the buffer
        mbar_wait_guard(bar(B_OUTREADY), j & 1);
        for (int kb = 0; kb < KB_X; ++kb) {
          if (p.inplace) tma_reduce_add_2d(&tmOut, sbase + OFF_X + (j & 1) * X_BYTES + kb * KBLK, kb * 64, tile_row(j));
          else tma_store_2d(&tmOut, sbase + OFF_X + (j & 1) * X_BYTES + kb * KBLK, kb * 64, tile_row(j));
        }
        bulk_commit();
        if (j + 2 < nt) { bulk_wait_read0(); load_x(j + 2); }
      }
      bulk_wait0();
    }
  } else if (warp == 2) {
    // =============================== W1 ring (this CTA's half of every K-block) ===============================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int g = 0; g < nt * NCHUNK; ++g) {
        const int c = g % NCHUNK;
        mbar_wait_guard(bar(B_W1EMPTY + s), ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(bar(B_W1FULL + s), 2 * S1_BYTES);       // bytes of BOTH CTAs land on the leader's barrier
        const uint32_t lbar = mapa_rank(bar(B_W1FULL + s), 0);
        for (int kb = 0; kb < KB_X; ++kb)
          tma_load_2d_2sm(sbase + OFF_W1 + s * S1_BYTES + kb * S1K, &tmW1, kb * 64, c * HC + (int)rank * (HC / 2), lbar);
        if (++s == NS1) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // =============================== W2 ring ===============================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int g = 0; g < nt * NCHUNK; ++g) {
        const int c = g % NCHUNK;
        mbar_wait_guard(bar(B_W2EMPTY + s), ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(bar(B_W2FULL + s), 2 * S2_BYTES);
        const uint32_t lbar = mapa_rank(bar(B_W2FULL + s), 0);
        for (int kb = 0; kb < KB_H; ++kb)
          tma_load_2d_2sm(sbase + OFF_W2 + s * S2_BYTES + kb * S2K, &tmW2, c * HC + kb * 64, (int)rank * (D / 2), lbar);
        if (++s == NS2) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (rank == 0 && nt > 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(2 * TM, HC, false, false);   // FC1: M = 256, N = 128
      constexpr uint32_t idesc2 = umma_idesc_bf16(2 * TM, D, false, false);    // FC2: M = 256, N = 192
      int s1 = 0, s2 = 0; uint32_t ph1 = 0, ph2 = 0;
      const int G = nt * NCHUNK;
      long long w_xn = 0, w_w1 = 0, w_a2 = 0, w_h = 0, w_w2 = 0;
      auto fc1 = [&](int g) {
        const int j = g / NCHUNK, c = g % NCHUNK;
        const uint32_t b = g & 1;
        if (c == 0) FM2_WAIT(w_xn, mbar_wait_guard(bar(B_XNREADY + (j & 1)), (j >> 1) & 1));   // LayerNorm of tile j done in both CTAs
        tc_fence_after();
        FM2_STAMP(100 + 4 * c);
        FM2_WAIT(w_w1, mbar_wait_guard(bar(B_W1FULL + s1), ph1));
        tc_fence_after();
        {
          const uint32_t la = umma_desc_lo(sbase + OFF_X + (j & 1) * X_BYTES), lb = umma_desc_lo(sbase + OFF_W1 + s1 * S1_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < KB_X; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(tmem_base + b * HC, umma_desc_from_lo(la + kb * (KBLK >> 4) + 2 * k), umma_desc_from_lo(lb + kb * (S1K >> 4) + 2 * k),
                              idesc1, (kb | k) ? 1u : 0u);
            umma_commit_2sm(bar(B_W1EMPTY + s1));
            umma_commit_2sm(bar(B_ACC1FULL + b));
          }
          __syncwarp();
          if (++s1 == NS1) { s1 = 0; ph1 ^= 1; }
        }
        FM2_STAMP(101 + 4 * c);
      };
      auto fc2 = [&](int g) {
        const int j = g / NCHUNK, c = g % NCHUNK;
        const uint32_t b = g & 1;
        if (c == 0) FM2_WAIT(w_a2, mbar_wait_guard(bar(B_ACC2EMPTY), (j & 1) ^ 1));   // previous tile's final epilogue drained acc2
        FM2_WAIT(w_h, mbar_wait_guard(bar(B_HREADY + b), (g >> 1) & 1));
        tc_fence_after();
        FM2_STAMP(102 + 4 * c);
        FM2_WAIT(w_w2, mbar_wait_guard(bar(B_W2FULL + s2), ph2));
        tc_fence_after();
        {
          const uint32_t lb = umma_desc_lo(sbase + OFF_W2 + s2 * S2_BYTES);
          const uint32_t ta = tmem_base + b * HC;                               // K-step kk -> columns (kk >> 1) * 32 + (kk & 1) * 8
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_2sm_ts(tmem_base + ACC2_COL, ta + (kk >> 1) * 32 + (kk & 1) * 8, umma_desc_from_lo(lb + (kk >> 2) * (S2K >> 4) + 2 * (kk & 3)),
                               idesc2, (c | kk) ? 1u : 0u);
            umma_commit_2sm(bar(B_W2EMPTY + s2));
            if (c == NCHUNK - 1) umma_commit_2sm(bar(B_ACC2FULL));
          }
          __syncwarp();
          if (++s2 == NS2) { s2 = 0; ph2 ^= 1; }
        }
        FM2_STAMP(103 + 4 * c);
      };
      fc1(0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) fc1(g + 1);
        fc2(g);
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[200] = w_xn; p.dbg[202] = w_w1; p.dbg[203] = w_a2; p.dbg[204] = w_h; p.dbg[205] = w_w2; }
    }
  } else {
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // row within this CTA's tile
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);
    if (warp < FIRST_AUX) {
      // =============================== GELU warps: thread = (row, 32 or 64 of the chunk's 128 columns) ===============================
      // The GELU epilogue is bound by the MUFU pipe (tanh: 16 lanes/clk/SM -> ~1000 cycles per chunk whatever the warp count), so it
      // gets its own warps and nothing else: LayerNorm and the final epilogue run concurrently on the aux warps.
      // NGW = 16 was measured: per launch it is 2.5 % FASTER in a short, un-throttled run (the MMA issuer's HREADY waits drop from
      // 2.4 k to 1.7 k cycles per tile) but 1.8 % SLOWER in the sustained rollout loop, where the GPU runs at its power cap
      // (~1740 MHz) and 256 more polling threads per SM cost more clock than the shorter waits return.  8 it is.
      constexpr int GROUPS = 4 / (NGW / 4);    // 32-column groups of a chunk per thread (NGW = 16 -> 1, NGW = 8 -> 2)
      const int cg = (warp - FIRST_CW) >> 2;   // column group set (0 .. NGW/4 - 1)
      const uint32_t l_hready = mapa_rank(bar(B_HREADY), 0);
      int g = 0;
      long long w_a1f = 0;
      for (int j = 0; j < nt; ++j) {
        const bool stamp = (warp == FIRST_CW && lane == 0);
        if (stamp) FM2_STAMP(0);
        for (int c = 0; c < NCHUNK; ++c, ++g) {
          const uint32_t b = g & 1, use = (g >> 1) & 1;
          FM2_WAIT(w_a1f, mbar_wait_guard(bar(B_ACC1FULL + b), use));
          tc_fence_after();
          if (stamp) FM2_STAMP(10 + 4 * c);
#pragma unroll
          for (int hh = 0; hh < GROUPS; ++hh) {
            const int c32 = cg * GROUPS + hh;                               // 32-column group of the chunk
            uint32_t r0[32];
            tmem_ld_32x32(tmem_base + tm_lane + b * HC + c32 * 32, r0);
            tmem_ld_wait();
            const uint4* bias = reinterpret_cast<const uint4*>(s_b1p + (c * HC + c32 * 32) / 2);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 bv = bias[i];
              pk[4 * i + 0] = gelu2_bf16x2(bf16x2_add(pack_bf16(__uint_as_float(r0[8 * i + 0]), __uint_as_float(r0[8 * i + 1])), bv.x));
              pk[4 * i + 1] = gelu2_bf16x2(bf16x2_add(pack_bf16(__uint_as_float(r0[8 * i + 2]), __uint_as_float(r0[8 * i + 3])), bv.y));
              pk[4 * i + 2] = gelu2_bf16x2(bf16x2_add(pack_bf16(__uint_as_float(r0[8 * i + 4]), __uint_as_float(r0[8 * i + 5])), bv.z));
              pk[4 * i + 3] = gelu2_bf16x2(bf16x2_add(pack_bf16(__uint_as_float(r0[8 * i + 6]), __uint_as_float(r0[8 * i + 7])), bv.w));
            }
            // H (packed bf16, two K=16 steps) over the first 16 columns of the 32-column group it was read from: the FC2 A operand
            tmem_st_32x16(tmem_base + tm_lane + b * HC + c32 * 32, pk);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(l_hready + 8u * b);
          if (stamp) FM2_STAMP(12 + 4 * c);
        }
      }
      if (p.dbg && blockIdx.x == 0 && warp == FIRST_CW && lane == 0) p.dbg[210] = w_a1f;
    } else {
      // =============================== aux warps: LayerNorm + final epilogue, thread = (row, 96 of the 192 columns) ===============================
      const int half = (warp - FIRST_AUX) >> 2;
      const int col0 = half * 96;
      const uint32_t l_xnready = mapa_rank(bar(B_XNREADY), 0), l_acc2empty = mapa_rank(bar(B_ACC2EMPTY), 0);
      const bool stamp = (warp == FIRST_AUX && lane == 0);

      // ---- LayerNorm (scale/shift folded into W1'/b1') in place in the x buffer ----
      auto layer_norm = [&](int jn) {
        uint8_t* xb = sptr + OFF_X + (jn & 1) * X_BYTES;
        mbar_wait_guard(bar(B_XFULL + (jn & 1)), (jn >> 1) & 1);
        uint4 v[12];
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int col = col0 + i * 8, kb = col >> 6, c8 = (col & 63) >> 3;
          v[i] = *reinterpret_cast<const uint4*>(xb + kb * KBLK + row * 128 + ((c8 ^ sw) << 4));
          stats_bf16x2(v[i].x, s, q); stats_bf16x2(v[i].y, s, q); stats_bf16x2(v[i].z, s, q); stats_bf16x2(v[i].w, s, q);
        }
        ln_part[row * 2 + half] = make_float2(s, q);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float2 o = ln_part[row * 2 + (half ^ 1)];
        s += o.x; q += o.y;
        const float mean = s * (1.0f / D);
        const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
        const uint32_t rstd_b = pack_bf16(rsqrtf(var + p.eps), 0.f);
        const float nmr = -mean * bf16_lo(rstd_b);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int col = col0 + i * 8, kb = col >> 6, c8 = (col & 63) >> 3;
          v[i].x = norm_bf16x2(v[i].x, rstd_b, nmr); v[i].y = norm_bf16x2(v[i].y, rstd_b, nmr);
          v[i].z = norm_bf16x2(v[i].z, rstd_b, nmr); v[i].w = norm_bf16x2(v[i].w, rstd_b, nmr);
          *reinterpret_cast<uint4*>(xb + kb * KBLK + row * 128 + ((c8 ^ sw) << 4)) = v[i];
        }
        fence_proxy_async_smem();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(l_xnready + 8u * (jn & 1));
        asm volatile("bar.sync 1, 256;" ::: "memory");   // ln_part reuse safety
      };

      // ---- final epilogue: out = acc2 + b2 + x.  Residual re-read from L2 (issued before the accumulator wait); the output
      //      tile is staged in the x buffer (LN(x) is dead once ACC2FULL has fired) and stored by warp 0 (TMA) ----
      auto final_epilogue = [&](int j) {
        mbar_wait_guard(bar(B_ACC2FULL), j & 1);
        tc_fence_after();
        if (stamp) FM2_STAMP(60);
        // drain acc2 first (packed to bf16 at once) so that the next tile's FC2(0) is released as early as possible ...
        uint32_t acc[48];
#pragma unroll
        for (int g3 = 0; g3 < 3; ++g3) {
          uint32_t ra[32];
          tmem_ld_32x32(tmem_base + tm_lane + ACC2_COL + col0 + g3 * 32, ra);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[g3 * 16 + i] = pack_bf16(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(l_acc2empty);
        // ... then add bias + residual (re-read from L2; these warps have slack) and stage the tile in the x buffer
        const int grow = tile_row(j) + row;
        uint8_t* xb = sptr + OFF_X + (j & 1) * X_BYTES;
        const uint4* gx = reinterpret_cast<const uint4*>(p.x + (size_t)grow * D + col0);
#pragma unroll
        for (int g3 = 0; g3 < 3; ++g3) {
          uint4 xr[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xr[i] = (!p.inplace && grow < p.M) ? gx[g3 * 4 + i] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = col0 + g3 * 32 + 8 * i, kb = col >> 6, c8 = (col & 63) >> 3;
            const uint4 bv = *reinterpret_cast<const uint4*>(s_b2 + (col >> 1));
            const uint32_t* a4 = &acc[g3 * 16 + 4 * i];
            uint4 ov;
            ov.x = bf16x2_add(bf16x2_add(a4[0], bv.x), xr[i].x);
            ov.y = bf16x2_add(bf16x2_add(a4[1], bv.y), xr[i].y);
            ov.z = bf16x2_add(bf16x2_add(a4[2], bv.z), xr[i].z);
            ov.w = bf16x2_add(bf16x2_add(a4[3], bv.w), xr[i].w);
            *reinterpret_cast<uint4*>(xb + kb * KBLK + row * 128 + ((((uint32_t)c8) ^ sw) << 4)) = ov;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_OUTREADY));
        if (stamp) FM2_STAMP(61);
      };

      // x(0), x(1) are loaded up front; x(j+2) is loaded into tile j's buffer once its output store has read it
      if (nt > 0) layer_norm(0);
      if (nt > 1) layer_norm(1);
      for (int j = 0; j < nt; ++j) {
        final_epilogue(j);
        if (j + 2 < nt) { layer_norm(j + 2); if (stamp) FM2_STAMP(50); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // neither CTA exits while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(fmlp2::TMEM_COLS) : "memory");
  }
}

bool fused_mlp2_supported(int D, int hidden) { return D == fmlp2::D && hidden == fmlp2::HID; }

// w1f: [HID, D] bf16 = W1 . diag(gamma);  b1p: [HID] bf16 = b1 + W1 . beta;  w2h: [D, HID] bf16 = W2 / 2  (vit_fold.cu)
int launch_fused_mlp2(cudaStream_t stream, const __nv_bfloat16* x, __nv_bfloat16* out, const __nv_bfloat16* w1f, const __nv_bfloat16* b1p,
                      const __nv_bfloat16* w2h, const float* b2, int M, int D, int hidden, float eps, const FusedOpts& o) {
  using namespace fmlp2;
  if (D != fmlp2::D || hidden != HID) { set_last_error("fused_mlp2: only D=192, hidden=768"); return VITMARL_EINVAL; }
  if (M <= 0) return VITMARL_OK;
  CUtensorMap tmX, tmW1, tmW2, tmOut;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmX, x, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmOut, out, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmW1, w1f, HID, D, (uint64_t)D * 2, HC / 2, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmW2, w2h, fmlp2::D, HID, (uint64_t)HID * 2, fmlp2::D / 2, 64))) return rc;
  FusedMlp2Params p{M, x, reinterpret_cast<const uint32_t*>(b1p), b2, eps, o.dbg, out == x ? 1 : 0};
  cudaError_t e = cudaFuncSetAttribute(fused_mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e);
  const int pair_tiles = (M + 2 * TM - 1) / (2 * TM);
  const int clusters = min(pair_tiles, num_sms() / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = o.pdl ? 1 : 0;
  return check_cuda(cudaLaunchKernelEx(&cfg, fused_mlp2_kernel, tmX, tmW1, tmW2, tmOut, p));
}

}  // namespace vitmarl
