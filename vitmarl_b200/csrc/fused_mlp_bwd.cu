// Fused BACKWARD of the transformer MLP block for D = 192 / hidden 768 (ViT-Tiny), training path:
//     forward (fused_mlp2.cu):  out = x + FC2( gelu_tanh( FC1( LN(x) ) ) )        saved: x only
//     this kernel, per 128-token tile, from (x, dY = d out):
//         xhat  = (x - mean) * rstd                                  -> global (B operand of the dW1' product)
//         hpre  = xhat . W1'^T + b1'           (recomputed, W1' = W1 diag(gamma), b1' = b1 + W1 beta: vit_fold.cu)
//         h2    = 2 gelu(hpre)                                       -> global (B operand of dW2 = 1/2 dY^T h2)
//         dhpre = (dY . W2h) * 2 gelu'(hpre)   (W2h = W2 / 2)        -> global (A operand of dW1' = dhpre^T xhat; column sums = db1')
//         dxhat = dhpre . W1'
//         dx    = dY + rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))     -> global
// Nothing of the forward's [tokens, 768] activations is saved: the MLP half of a training step moves 24 instead of 42
// [tokens, 192]-sized units through HBM.  The two weight-gradient products stay separate launches of gemm2_dw_kernel (a dW
// accumulator for one 128-wide hidden chunk alone is 2 x 192 TMEM columns); LayerNorm-parameter and unfolded-weight gradients
// follow from dW1' / db1' in mlp_bwd_unfold_kernel below.
//
// The kernel is HBM-bound on writing h2 and dhpre (590 KB per tile against ~14 k tensor cycles), so the structure is kept
// simple: one CTA per SM, 128-token tiles, hidden dimension in 12 chunks of 64:
//     FC1(c):  acc1      = xhat[tmem] . W1'[c]^T            (TS, N = 64, K = 192; LayerNorm writes xhat straight into tensor memory)
//     G(c):    accD[c&1] = dY[smem] . W2h[:, c]             (SS, N = 64, K = 192, B read MN-major)
//     epilogue: h2, dhpre from (acc1, accD); dhpre -> TMEM over accD (bf16 pairs) and, with h2, -> global through two staging tiles
//     X(c):    acc3 += dhpre[tmem] . W1'[c]                 (TS, N = 192, K = 64; the SAME W1' stage read MN-major)
// issue order FC1(c+1), G(c+1), X(c): the tensor pipe works on the next chunk while the epilogue warps run the GELU maths.
// Shared memory (216 KB): the x tile is dead once xhat is in tensor memory and stored, so its 48 KB become the two chunk staging
// tiles; that leaves room for a 3-deep W1' ring (a stage lives from its load to X(c): ~7 k cycles against ~2 k per chunk) next to
// a 2-deep W2h ring (a stage only lives until G(c) retires).  (First version: x as a shared-memory operand, ONE staging tile and a
// 2-deep combined ring -- 0.98 ms per launch at 4096 tiles, bound by the epilogue warps waiting on the staging hand-over and by
// the ring's chain latency.)
// Warp roles (640 threads): 0 x / dY tile loads, 1 TMEM allocator + MMA issuer, 2 weight rings, 3 TMA stores, 4-19 epilogue
// (thread = (token row, quarter)): LayerNorm, chunk epilogue on 16 of the 64 chunk columns, LayerNorm backward on 48 of 192 columns.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace fmb {
constexpr int D = 192, HID = 768, HC = 64, NCHUNK = HID / HC, TM = 128;
constexpr int KB_X = D / 64;                       // 3 K-blocks of the token operands
constexpr int KBLK = TM * 128;                     // [128 x 64] bf16 = 16 KB
constexpr int X_BYTES = KB_X * KBLK;               // 48 KB
constexpr int W1_BYTES = HC * D * 2;               // [64 x 192] as 3 K-blocks of [64 x 64] = 24 KB
constexpr int W2_BYTES = D * HC * 2;               // [192 x 64] = 24 KB
constexpr int NS1 = 3, NS2 = 2;                    // W1' stages live from FC1(c) to X(c) (load + FC1 + epilogue + X ~ 7 k cycles): 3 deep;
                                                   // W2h stages only until G(c) has retired: 2 deep
constexpr int OFF_X = 0;                           // x landing tile -> xhat (in place; copied to TMEM and stored) -> the two chunk staging tiles
constexpr int OFF_STG = OFF_X;                     // h2 tile @ +0, dhpre tile @ +16 KB  (aliases the dead x tile during the chunk loop)
constexpr int OFF_DY = OFF_X + X_BYTES;            // dY tile (A operand of G) -> dx staging in place
constexpr int OFF_W1 = OFF_DY + X_BYTES;
constexpr int OFF_W2 = OFF_W1 + NS1 * W1_BYTES;
constexpr int OFF_BAR = OFF_W2 + NS2 * W2_BYTES;
constexpr int OFF_MISC = OFF_BAR + 256;
constexpr int MISC_BYTES = 128 * 4 * 8 + HID * 2 + (HID + D) * 4;  // row partials [128][4] float2, b1' (bf16), column sums of dhpre [HID] and dx [D] (fp32)
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int W_LOAD = 0, W_MMA = 1, W_WRING = 2, W_STORE = 3, W_E0 = 4, N_E = 16;
constexpr int THREADS = 32 * (W_E0 + N_E);         // 640
constexpr int TMEM_COLS = 512;
constexpr int COL_XN = 0, COL_A1 = 96, COL_AD = 160, COL_A3 = 288;   // xhat (bf16 pairs) 0..95 | acc1 96..159 | accD[2] 160..287 | acc3 288..479
enum { B_XFULL = 0, B_DYFULL, B_XNREADY, B_XSTORED, B_XFREE, B_DYFREE, B_OUTREADY, B_ACC3FULL, B_ACC3FREE, B_A1FREE,
       B_STGFULL, B_STGFREE = B_STGFULL + 2, B_ACCFULL = B_STGFREE + 2, B_HREADY = B_ACCFULL + 2,
       B_W1FULL = B_HREADY + 2, B_W1EMPTY = B_W1FULL + NS1, B_W2FULL = B_W1EMPTY + NS1, B_W2EMPTY = B_W2FULL + NS2, B_TMEMSLOT = B_W2EMPTY + NS2, B_COUNT };
static_assert(B_COUNT * 8 <= 256, "barrier area");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace fmb

struct FusedMlpBwdParams {
  int M;
  const uint16_t* b1p;          // [HID] folded FC1 bias, bf16
  float eps;
  float* dbf;                   // [HID] += column sums of dhpre  (gradient of the folded FC1 bias; nullable)
  float* dbx;                   // [D]   += column sums of dx     (the bias gradient of whatever produced this block's input; nullable)
  long long* dbg;               // optional clock64 timeline of CTA 0, its second tile (null in production)
};
#define FMB_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && j == 1 && lane == 0) p.dbg[(slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(fmb::THREADS, 1)
fused_mlp_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmXH, const __grid_constant__ CUtensorMap tmH,
                     const __grid_constant__ CUtensorMap tmDH, const __grid_constant__ CUtensorMap tmDX, const FusedMlpBwdParams p) {
  using namespace fmb;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  float2* part = reinterpret_cast<float2*>(sptr + OFF_MISC);                       // [128][4]
  uint16_t* s_b1 = reinterpret_cast<uint16_t*>(sptr + OFF_MISC + 128 * 4 * 8);     // [HID] bf16
  float* s_cs = reinterpret_cast<float*>(sptr + OFF_MISC + 128 * 4 * 8 + HID * 2);  // [HID + D] per-CTA column sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M + TM - 1) / TM;
  const int nt = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_row = [&](int j) { return ((int)blockIdx.x + j * (int)gridDim.x) * TM; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmXH); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmDH); tma_prefetch_desc(&tmDX);
    mbar_init(bar(B_XFULL), 1); mbar_init(bar(B_DYFULL), 1); mbar_init(bar(B_XNREADY), N_E); mbar_init(bar(B_XSTORED), 1); mbar_init(bar(B_XFREE), 1);
    mbar_init(bar(B_DYFREE), 1); mbar_init(bar(B_OUTREADY), N_E); mbar_init(bar(B_ACC3FULL), 1); mbar_init(bar(B_ACC3FREE), N_E);
    mbar_init(bar(B_A1FREE), N_E);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_STGFULL + i), N_E); mbar_init(bar(B_STGFREE + i), 1); mbar_init(bar(B_ACCFULL + i), 1); mbar_init(bar(B_HREADY + i), N_E);
    }
    for (int i = 0; i < NS1; ++i) { mbar_init(bar(B_W1FULL + i), 1); mbar_init(bar(B_W1EMPTY + i), 1); }
    for (int i = 0; i < NS2; ++i) { mbar_init(bar(B_W2FULL + i), 1); mbar_init(bar(B_W2EMPTY + i), 1); }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(bar(B_TMEMSLOT));
  for (int i = threadIdx.x; i < HID; i += THREADS) s_b1[i] = p.b1p[i];
  for (int i = threadIdx.x; i < HID + D; i += THREADS) s_cs[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMSLOT)));

  if (warp == W_LOAD) {
    // =============================== x / dY tiles ===============================
    if (lane == 0) {
      for (int j = 0; j < nt; ++j) {
        mbar_wait_guard(bar(B_XFREE), (j & 1) ^ 1);            // the last chunk stores of tile j-1 have read the staging tiles (= this buffer)
        mbar_arrive_expect_tx(bar(B_XFULL), X_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_X + kb * KBLK, &tmX, kb * 64, tile_row(j), bar(B_XFULL));
        mbar_wait_guard(bar(B_DYFREE), (j & 1) ^ 1);           // the dx store of tile j-1 has read the dY buffer
        mbar_arrive_expect_tx(bar(B_DYFULL), X_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_DY + kb * KBLK, &tmDY, kb * 64, tile_row(j), bar(B_DYFULL));
      }
    }
  } else if (warp == W_WRING) {
    // =============================== weight rings: W1'[c] (3 K-blocks of [64 x 64]) and W2h[:, c] ([192 x 64]) ===============================
    if (lane == 0) {
      int s1 = 0, s2 = 0; uint32_t ph1 = 0, ph2 = 0;
      for (int g = 0; g < nt * NCHUNK; ++g) {
        const int c = g % NCHUNK;
        mbar_wait_guard(bar(B_W2EMPTY + s2), ph2 ^ 1);
        mbar_arrive_expect_tx(bar(B_W2FULL + s2), W2_BYTES);
        tma_load_2d(sbase + OFF_W2 + s2 * W2_BYTES, &tmW2, c * HC, 0, bar(B_W2FULL + s2));
        if (++s2 == NS2) { s2 = 0; ph2 ^= 1; }
        mbar_wait_guard(bar(B_W1EMPTY + s1), ph1 ^ 1);
        mbar_arrive_expect_tx(bar(B_W1FULL + s1), W1_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_W1 + s1 * W1_BYTES + kb * (HC * 128), &tmW1, kb * 64, c * HC, bar(B_W1FULL + s1));
        if (++s1 == NS1) { s1 = 0; ph1 ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer ===============================
    if (nt > 0) {
      constexpr uint32_t id_fc1 = umma_idesc_bf16(TM, HC, false, false);     // A = xhat in tensor memory
      constexpr uint32_t id_g = umma_idesc_bf16(TM, HC, false, true);        // B = W2h[:, c] read MN-major (rows = output features of FC2)
      constexpr uint32_t id_x = umma_idesc_bf16(TM, D, false, true);         // B = W1'[c] read MN-major (rows = hidden units of the chunk)
      int s1f = 0, s2f = 0; uint32_t ph1f = 0, ph2f = 0;   // ring positions of the next FC1 / G
      int s1x = 0;                                          // ring position of the next X product
      uint32_t g_f = 0;                                     // global chunk counter of FC1 / G
      auto fc1g = [&](int j) {
        const uint32_t b = g_f & 1;
        const int cc = (int)(g_f % NCHUNK);
        FMB_STAMP(100 + 6 * cc);
        // acc1 is single-buffered: the epilogue of the previous chunk drains it first thing
        mbar_wait_guard(bar(B_A1FREE), (g_f & 1) ^ 1);
        FMB_STAMP(101 + 6 * cc);
        mbar_wait_guard(bar(B_W1FULL + s1f), ph1f);
        FMB_STAMP(102 + 6 * cc);
        tc_fence_after();
        const uint32_t lb1 = umma_desc_lo(sbase + OFF_W1 + s1f * W1_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < KB_X; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tmem_base + COL_A1, tmem_base + COL_XN + (kb * 4 + k) * 8, umma_desc_from_lo(lb1 + kb * ((HC * 128) >> 4) + 2 * k), id_fc1,
                           (kb | k) ? 1u : 0u);
        }
        __syncwarp();
        mbar_wait_guard(bar(B_W2FULL + s2f), ph2f);
        FMB_STAMP(103 + 6 * cc);
        tc_fence_after();
        const uint32_t lay = umma_desc_lo(sbase + OFF_DY), lb2 = umma_desc_lo(sbase + OFF_W2 + s2f * W2_BYTES, 8192);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < KB_X; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + COL_AD + b * HC, umma_desc_from_lo(lay + kb * (KBLK >> 4) + 2 * k), umma_desc_from_lo(lb2 + (kb * 4 + k) * 128),
                        id_g, (kb | k) ? 1u : 0u);
          umma_commit(bar(B_W2EMPTY + s2f));                   // W2h[:, c] is dead once G(c) has retired
          umma_commit(bar(B_ACCFULL + b));
        }
        __syncwarp();
        if (++s1f == NS1) { s1f = 0; ph1f ^= 1; }
        if (++s2f == NS2) { s2f = 0; ph2f ^= 1; }
        ++g_f;
      };
      uint32_t g_x = 0;
      for (int j = 0; j < nt; ++j) {
        mbar_wait_guard(bar(B_XNREADY), j & 1);               // xhat of tile j is in tensor memory
        mbar_wait_guard(bar(B_DYFULL), j & 1);
        tc_fence_after();
        fc1g(j);
        for (int c = 0; c < NCHUNK; ++c, ++g_x) {
          if (c + 1 < NCHUNK) fc1g(j);                        // next chunk's products run under this chunk's epilogue
          const uint32_t b = g_x & 1;
          FMB_STAMP(104 + 6 * c);
          mbar_wait_guard(bar(B_HREADY + b), (g_x >> 1) & 1);
          FMB_STAMP(105 + 6 * c);
          if (c == 0) mbar_wait_guard(bar(B_ACC3FREE), (j & 1) ^ 1);     // previous tile's LayerNorm-backward epilogue drained acc3
          tc_fence_after();
          const uint32_t lb = umma_desc_lo(sbase + OFF_W1 + s1x * W1_BYTES, 8192);
          const uint32_t ta = tmem_base + COL_AD + b * HC;     // dhpre packed: K step kk (16 hidden units) -> the first 8 columns of quarter kk
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(tmem_base + COL_A3, ta + kk * 16, umma_desc_from_lo(lb + kk * 128), id_x, (c | kk) ? 1u : 0u);
            umma_commit(bar(B_W1EMPTY + s1x));
            if (c == NCHUNK - 1) umma_commit(bar(B_ACC3FULL));
          }
          __syncwarp();
          if (++s1x == NS1) s1x = 0;
        }
      }
    }
  } else if (warp == W_STORE) {
    // =============================== TMA stores ===============================
    if (lane == 0) {
      uint32_t g = 0;
      for (int j = 0; j < nt; ++j) {
        mbar_wait_guard(bar(B_XNREADY), j & 1);
        for (int kb = 0; kb < KB_X; ++kb) tma_store_2d(&tmXH, sbase + OFF_X + kb * KBLK, kb * 64, tile_row(j));
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(B_XSTORED));                          // the x tile is dead: its memory becomes the two staging tiles
        // both staging tiles of a chunk are handed over together: two stores in one bulk group, handed back once both are read
        for (int c = 0; c < NCHUNK; ++c, ++g) {
          mbar_wait_guard(bar(B_STGFULL + 0), g & 1);
          mbar_wait_guard(bar(B_STGFULL + 1), g & 1);
          tma_store_2d(&tmH, sbase + OFF_STG, c * HC, tile_row(j));
          tma_store_2d(&tmDH, sbase + OFF_STG + TM * HC * 2, c * HC, tile_row(j));
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(bar(B_STGFREE + 0));
          mbar_arrive(bar(B_STGFREE + 1));
        }
        mbar_arrive(bar(B_XFREE));                             // staging tiles read: the next tile's x may land
        mbar_wait_guard(bar(B_OUTREADY), j & 1);
        for (int kb = 0; kb < KB_X; ++kb) tma_store_2d(&tmDX, sbase + OFF_DY + kb * KBLK, kb * 64, tile_row(j));
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(B_DYFREE));
      }
      bulk_wait0();
    }
  } else if (warp >= W_E0) {
    // =============================== epilogue warps: thread = (token row, quarter) ===============================
    // 16 warps (4 per scheduler): the chunk epilogue is ~800 instructions per 32 elements, and with 8 warps it ran at ~0.6 IPC
    // per scheduler (2 750 cycles per chunk, the kernel's critical path)
    const int quad = warp & 3, qt = (warp - W_E0) >> 2;       // TMEM lane quadrant; column quarter
    const int row = quad * 32 + lane;
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);
    const int col0 = qt * 48;                                  // LayerNorm / LayerNorm-backward columns of this thread
    uint32_t g = 0;
    for (int j = 0; j < nt; ++j) {
      // ---- LayerNorm (scale / shift folded into W1' / b1'): x -> xhat in place (for the store) and into tensor memory (FC1's A operand) ----
      float rstd;
      {
        uint8_t* xb = sptr + OFF_X;
        if (warp == W_E0) FMB_STAMP(0);
        mbar_wait_guard(bar(B_XFULL), j & 1);
        if (warp == W_E0) FMB_STAMP(1);
        uint4 v[6];
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int col = col0 + i * 8, kb = col >> 6, c8 = (col & 63) >> 3;
          v[i] = *reinterpret_cast<const uint4*>(xb + kb * KBLK + row * 128 + ((c8 ^ sw) << 4));
          stats_bf16x2(v[i].x, s, q); stats_bf16x2(v[i].y, s, q); stats_bf16x2(v[i].z, s, q); stats_bf16x2(v[i].w, s, q);
        }
        part[row * 4 + qt] = make_float2(s, q);
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int k = 1; k < 4; ++k) { const float2 o = part[row * 4 + ((qt + k) & 3)]; s += o.x; q += o.y; }
        const float mean = s * (1.0f / D);
        const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
        const uint32_t rstd_b = pack_bf16(rsqrtf(var + p.eps), 0.f);      // (the forward normalises with the bf16-rounded rstd)
        rstd = bf16_lo(rstd_b);
        const float nmr = -mean * rstd;
        uint32_t w[24];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int col = col0 + i * 8, kb = col >> 6, c8 = (col & 63) >> 3;
          v[i].x = norm_bf16x2(v[i].x, rstd_b, nmr); v[i].y = norm_bf16x2(v[i].y, rstd_b, nmr);
          v[i].z = norm_bf16x2(v[i].z, rstd_b, nmr); v[i].w = norm_bf16x2(v[i].w, rstd_b, nmr);
          *reinterpret_cast<uint4*>(xb + kb * KBLK + row * 128 + ((c8 ^ sw) << 4)) = v[i];
          w[4 * i] = v[i].x; w[4 * i + 1] = v[i].y; w[4 * i + 2] = v[i].z; w[4 * i + 3] = v[i].w;
        }
        // xhat as bf16 pairs: columns qt * 24 .. + 24 of XN (the previous tile's FC1 products retired before its ACC3FULL)
#pragma unroll
        for (int t3 = 0; t3 < 3; ++t3) {
          uint32_t w8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w8[i] = w[t3 * 8 + i];
          tmem_st_32x8(tmem_base + tm_lane + COL_XN + qt * 24 + t3 * 8, w8);
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_XNREADY));
        asm volatile("bar.sync 1, 512;" ::: "memory");        // `part` reuse safety
        if (warp == W_E0) FMB_STAMP(2);
      }
      // ---- chunk epilogues: thread = (row, 16 of the chunk's 64 hidden units) ----
      for (int c = 0; c < NCHUNK; ++c, ++g) {
        const uint32_t b = g & 1;
        if (warp == W_E0) FMB_STAMP(10 + 4 * c);
        mbar_wait_guard(bar(B_ACCFULL + b), (g >> 1) & 1);
        if (warp == W_E0) FMB_STAMP(11 + 4 * c);
        tc_fence_after();
        uint32_t a1[16], ad[16];
        tmem_ld_32x16(tmem_base + tm_lane + COL_A1 + qt * 16, a1);
        tmem_ld_32x16(tmem_base + tm_lane + COL_AD + b * HC + qt * 16, ad);
        tmem_ld_wait();
        if (warp == W_E0) FMB_STAMP(200 + 8 * c);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_A1FREE));            // acc1 drained: FC1 of the next chunk may overwrite it
        const uint32_t* bias = reinterpret_cast<const uint32_t*>(s_b1 + c * HC + qt * 16);
        uint32_t hw[8], dw[8];
        float dcs[16];                                        // this row's dhpre (fp32) for the bias-gradient column sums
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t bw = bias[i];
          float h[2], d[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float x = __uint_as_float(a1[2 * i + e]) + (e ? bf16_hi(bw) : bf16_lo(bw));
            const float k0 = 0.7978845608028654f, k1 = 0.044715f;
            const float x2 = x * x;
            const float t = tanh_approx(x * fmaf(k0 * k1, x2, k0));
            const float opt = 1.0f + t;
            h[e] = x * opt;                                                       // 2 gelu(x)
            const float g2 = fmaf(x * fmaf(-t, t, 1.0f), fmaf(3.0f * k0 * k1, x2, k0), opt);   // 2 gelu'(x)
            d[e] = __uint_as_float(ad[2 * i + e]) * g2;
          }
          hw[i] = pack_bf16(h[0], h[1]);
          dw[i] = pack_bf16(d[0], d[1]);
          dcs[2 * i] = d[0]; dcs[2 * i + 1] = d[1];
        }
        // dhpre (bf16 pairs) over the first 8 of this thread's own 16 accD columns: K step `qt` of the A operand of X(c)
        if (warp == W_E0) FMB_STAMP(201 + 8 * c);
        tc_fence_after();
        tmem_st_32x8(tmem_base + tm_lane + COL_AD + b * HC + qt * 16, dw);
        tmem_st_wait();
        if (warp == W_E0) FMB_STAMP(202 + 8 * c);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_HREADY + b));
        if (warp == W_E0) FMB_STAMP(12 + 4 * c);
        // h2 and dhpre to global through their staging tiles (the store warp drains them while the next chunk is computed).
        // The tiles alias the x tile: before the first write of a tile the xhat store must have read it.
        if (c == 0) { mbar_wait_guard(bar(B_XSTORED), j & 1); if (warp == W_E0) FMB_STAMP(3); }
        mbar_wait_guard(bar(B_STGFREE + 0), (g & 1) ^ 1);
        mbar_wait_guard(bar(B_STGFREE + 1), (g & 1) ^ 1);
        if (warp == W_E0) FMB_STAMP(203 + 8 * c);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          uint8_t* st = sptr + OFF_STG + which * (TM * HC * 2) + row * 128;
          const uint32_t* src = which ? dw : hw;
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2)
            *reinterpret_cast<uint4*>(st + ((((uint32_t)(qt * 2 + q2)) ^ sw) << 4)) = make_uint4(src[4 * q2], src[4 * q2 + 1], src[4 * q2 + 2], src[4 * q2 + 3]);
        }
        if (warp == W_E0) FMB_STAMP(204 + 8 * c);
        fence_proxy_async_smem();                             // one fence for both tiles
        if (warp == W_E0) FMB_STAMP(205 + 8 * c);
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar(B_STGFULL + 0)); mbar_arrive(bar(B_STGFULL + 1)); }
        // db1' : column sums of dhpre over the warp's 32 rows (transposing butterfly), accumulated per CTA in shared memory
        if (p.dbf) {
          const float cs = warp_colsum16(dcs, lane);
          if (lane < 16) atomicAdd(&s_cs[c * HC + qt * 16 + lane], cs);
        }
        if (warp == W_E0) FMB_STAMP(13 + 4 * c);
      }
      // ---- LayerNorm backward: dx = dY + rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), staged over dY in place ----
      {
        if (warp == W_E0) FMB_STAMP(70);
        mbar_wait_guard(bar(B_ACC3FULL), j & 1);
        if (warp == W_E0) FMB_STAMP(71);
        tc_fence_after();
        uint8_t* yb = sptr + OFF_DY;
        uint32_t xh[24];                                      // this thread's xhat (bf16 pairs) back from tensor memory
#pragma unroll
        for (int t3 = 0; t3 < 3; ++t3) {
          uint32_t w8[8];
          tmem_ld_32x8(tmem_base + tm_lane + COL_XN + qt * 24 + t3 * 8, w8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) xh[t3 * 8 + i] = w8[i];
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int g3 = 0; g3 < 3; ++g3) {
          uint32_t r[16];
          tmem_ld_32x16(tmem_base + tm_lane + COL_A3 + col0 + g3 * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d0 = __uint_as_float(r[2 * i]), d1 = __uint_as_float(r[2 * i + 1]);
            const uint32_t xw = xh[g3 * 8 + i];
            s1 += d0 + d1;
            s2 = fmaf(d0, bf16_lo(xw), fmaf(d1, bf16_hi(xw), s2));
          }
        }
        part[row * 4 + qt] = make_float2(s1, s2);
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int k = 1; k < 4; ++k) { const float2 o = part[row * 4 + ((qt + k) & 3)]; s1 += o.x; s2 += o.y; }
        const float m1 = s1 * (1.0f / D), m2 = s2 * (1.0f / D);
#pragma unroll
        for (int g3 = 0; g3 < 3; ++g3) {
          uint32_t r[16];
          tmem_ld_32x16(tmem_base + tm_lane + COL_A3 + col0 + g3 * 16, r);
          tmem_ld_wait();
          if (g3 == 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_ACC3FREE));      // accumulator (and XN) drained: the next tile may overwrite them
          }
          float ocs[16];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int col = col0 + g3 * 16 + 8 * i, kb = col >> 6, c8 = (col & 63) >> 3;
            const uint32_t off = kb * KBLK + row * 128 + ((((uint32_t)c8) ^ sw) << 4);
            const uint4 yv = *reinterpret_cast<const uint4*>(yb + off);
            const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
            uint32_t ow[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t xw = xh[g3 * 8 + 4 * i + e];
              const float d0 = __uint_as_float(r[8 * i + 2 * e]), d1 = __uint_as_float(r[8 * i + 2 * e + 1]);
              const float o0 = fmaf(rstd, d0 - m1 - bf16_lo(xw) * m2, bf16_lo(yw[e]));
              const float o1 = fmaf(rstd, d1 - m1 - bf16_hi(xw) * m2, bf16_hi(yw[e]));
              ow[e] = pack_bf16(o0, o1);
              ocs[8 * i + 2 * e] = o0; ocs[8 * i + 2 * e + 1] = o1;
            }
            *reinterpret_cast<uint4*>(yb + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);     // dx staged over dY, in place
          }
          if (p.dbx) {
            const float cs = warp_colsum16(ocs, lane);
            if (lane < 16) atomicAdd(&s_cs[HID + col0 + g3 * 16 + lane], cs);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_OUTREADY));
        if (warp == W_E0) FMB_STAMP(72);
        asm volatile("bar.sync 1, 512;" ::: "memory");        // `part` reuse safety
      }
    }
    // per-CTA column sums -> global (one fp32 reduction per column per CTA)
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (nt > 0) {
      const int t = threadIdx.x - W_E0 * 32;
      if (p.dbf) for (int i = t; i < HID; i += N_E * 32) atomicAdd(p.dbf + i, s_cs[i]);
      if (p.dbx) for (int i = t; i < D; i += N_E * 32) atomicAdd(p.dbx + i, s_cs[HID + i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc<fmb::TMEM_COLS>(tmem_base);
  }
}

// ---- unfolding the gradients of the folded parameters (vit_fold.cu): W' = c W diag(gamma), b' = c (b + W beta) ----
//     dW[o, i] += c (gamma[i] dW'[o, i] + beta[i] db'[o]);   dgamma[i] += c sum_o dW'[o, i] W[o, i];   dbeta[i] += c sum_o db'[o] W[o, i];
//     db[o] += c db'[o]                      (b' depends on W through W beta, hence the beta[i] db'[o] term of dW)
// One CTA per 8-row slab of W (rows o), 256 threads: thread = column; column sums through one global reduction per CTA and column.
__global__ void __launch_bounds__(256) unfold_grads_kernel(int N, int K, float c, const __nv_bfloat16* __restrict__ W, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ dWf, const float* __restrict__ dbf, float* __restrict__ dW,
                                                            float* __restrict__ db, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float red[];   // [2][K]
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  constexpr int ROWS = 8;
  const int o0 = blockIdx.x * ROWS;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const float gm = gamma ? gamma[i] : 1.0f, bt = (beta && dbf) ? beta[i] : 0.0f;
    float sg = 0.f, sb = 0.f;
    for (int o = o0; o < min(o0 + ROWS, N); ++o) {
      const float w = __bfloat162float(W[(size_t)o * K + i]);
      const float g = dWf[(size_t)o * K + i];
      dW[(size_t)o * K + i] += c * (gm * g + (dbf ? bt * dbf[o] : 0.0f));
      sg = fmaf(g, w, sg);
      if (dbf) sb = fmaf(dbf[o], w, sb);
    }
    red[i] = c * sg; red[K + i] = c * sb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta && dbf) atomicAdd(dbeta + i, red[K + i]);
  }
  if (db && dbf)
    for (int o = o0 + threadIdx.x; o < min(o0 + ROWS, N); o += blockDim.x) db[o] += c * dbf[o];
}

int launch_unfold_grads(cudaStream_t s, int N, int K, float c, const __nv_bfloat16* W, const float* gamma, const float* beta, const float* dWf,
                        const float* dbf, float* dW, float* db, float* dgamma, float* dbeta) {
  if (N <= 0 || K <= 0) return VITMARL_OK;
  unfold_grads_kernel<<<(N + 7) / 8, 256, 2 * K * sizeof(float), s>>>(N, K, c, W, gamma, beta, dWf, dbf, dW, db, dgamma, dbeta);
  return check_cuda(cudaGetLastError());
}

bool fused_mlp_bwd_supported(int D, int hidden) { return D == fmb::D && hidden == fmb::HID; }

// x, dy [M, D] bf16;  w1f [HID, D] bf16 folded, b1p [HID] bf16 folded, w2h [D, HID] bf16 = W2 / 2  (vit_fold.cu)
// outputs: xhat [M, D], h2 [M, HID] (= 2 gelu(hpre)), dh [M, HID] (= d hpre), dx [M, D]   (dx may alias dy: each tile's dy is read before its dx is written)
// dbf [HID] += column sums of dh, dbx [D] += column sums of dx (fp32, nullable): the two bias gradients that would otherwise each cost
// a pass over a [M, .] gradient
int launch_fused_mlp_bwd(cudaStream_t stream, const __nv_bfloat16* x, const __nv_bfloat16* dy, const __nv_bfloat16* w1f, const __nv_bfloat16* b1p,
                         const __nv_bfloat16* w2h, __nv_bfloat16* xhat, __nv_bfloat16* h2, __nv_bfloat16* dh, __nv_bfloat16* dx, int M, int D,
                         int hidden, float eps, float* dbf, float* dbx, long long* dbg) {
  using namespace fmb;
  if (D != fmb::D || hidden != HID) { set_last_error("fused_mlp_bwd: only D=192, hidden=768"); return VITMARL_EINVAL; }
  if (M <= 0) return VITMARL_OK;
  CUtensorMap tmX, tmDY, tmW1, tmW2, tmXH, tmH, tmDH, tmDX;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmX, x, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmDY, dy, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmXH, xhat, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmDX, dx, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmH, h2, M, HID, (uint64_t)HID * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmDH, dh, M, HID, (uint64_t)HID * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmW1, w1f, HID, D, (uint64_t)D * 2, HC, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmW2, w2h, fmb::D, HID, (uint64_t)HID * 2, fmb::D, 64))) return rc;
  FusedMlpBwdParams p{M, reinterpret_cast<const uint16_t*>(b1p), eps, dbf, dbx, dbg};
  cudaError_t e = cudaFuncSetAttribute(fused_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e);
  const int tiles = (M + TM - 1) / TM;
  fused_mlp_bwd_kernel<<<min(tiles, num_sms()), THREADS, SMEM_BYTES, stream>>>(tmX, tmDY, tmW1, tmW2, tmXH, tmH, tmDH, tmDX, p);
  return check_cuda(cudaGetLastError());
}

}  // namespace vitmarl
