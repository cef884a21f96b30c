// Fused attention block for D = 192, 3 heads x 64, T = 64 tokens (ViT-Tiny), inference path:
//     out = x + Wo . concat_h( softmax(Q_h K_h^T / 8) V_h ) + bo,    [Q|K|V] = LayerNorm(x) . Wqkv^T + bqkv
// ONE kernel per layer: the qkv / attention tensors ([tokens, 3D] + [tokens, D]) never touch HBM, and the
// LayerNorm pass, two GEMM launches and the attention launch of the unfused v0 collapse into it.
//
// Per CTA: a 128-token tile = two images (persistent over tiles).  Everything runs on tcgen05:
//   QKV_h  = LN(x)[128x192] . [Wq_h;Wk_h;Wv_h]^T        N = 192  (accumulator: TMEM cols   0..191)
//   S      = Q_h[128x64] . K_h[128x64]^T                 N = 128  (TMEM cols 192..319; the two images' scores are the
//                                                                  diagonal 64x64 blocks, the off-diagonal half is wasted)
//   O_h    = P[128x128] . V_h[128x64]                    N = 64, V_h read as an MN-major B operand (TMEM cols 320..383)
//   out    = concat(O_h)[128x192] . Wo^T                 N = 192  (reuses TMEM cols 0..191)
// Warp roles: warp 0 TMA producer (x tile + weight K-blocks), warp 1 TMEM alloc + MMA issuer, warps 2-9 compute
// (LayerNorm in smem, bias/convert epilogues into the swizzled smem operand buffers, softmax, final epilogue).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace fattn {
constexpr int D = 192, NH = 3, DH = 64, TM = 128;
constexpr int KB_X = D / 64;                       // 3 K-blocks
constexpr int NSTW = 3;                            // weight ring stages ([192 x 64] K-blocks, 24 KB)
constexpr int STAGE_BYTES = 192 * 128;             // 24 KB
constexpr int XN_BYTES = KB_X * TM * 128;          // 48 KB
constexpr int KBLK = TM * 128;                     // one [128 x 64] bf16 K-block = 16 KB
constexpr int OFF_XN = 0;
constexpr int OFF_QK = OFF_XN + XN_BYTES;          // Q (16 KB) | K (16 KB); reused as P (2 K-blocks of 64 keys)
constexpr int OFF_V = OFF_QK + 2 * KBLK;           // V_h [128 keys x 64 d]
constexpr int OFF_OC = OFF_V + KBLK;               // concat(O_h) [128 x 192] = 3 K-blocks
constexpr int OFF_W = OFF_OC + 3 * KBLK;
constexpr int OFF_BAR = OFF_W + NSTW * STAGE_BYTES;
constexpr int OFF_MISC = OFF_BAR + 256;
constexpr int MISC_BYTES = 128 * 2 * 8 + (2 * D + 3 * D + D) * 4;   // LN / softmax partials, gamma, beta, bqkv, bo
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int THREADS = 64 + 256;
constexpr int TMEM_COLS = 512;
constexpr int COL_QKV = 0, COL_S = 192, COL_O = 320;
enum { B_XFULL = 0, B_XEMPTY, B_XNREADY, B_QKVFULL, B_QKVEMPTY, B_QKREADY, B_SFULL, B_PREADY, B_OFULL, B_OEMPTY, B_PFULL, B_PEMPTY,
       B_RESFULL = 12, B_RESGO = 13, B_WFULL = 14, B_WEMPTY = B_WFULL + NSTW, B_TMEMSLOT = B_WEMPTY + NSTW, B_COUNT };
static_assert(B_COUNT * 8 <= 256, "barrier area");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace fattn

struct FusedAttnParams {
  int M;
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  const float* gamma; const float* beta;
  const float* bqkv; const float* bo;
  float eps;
  long long* dbg;               // optional clock64 timeline (CTA 0, second tile); null in production
};

#define FATT_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it == 1) p.dbg[(slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(fattn::THREADS, 1)
fused_attn_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWqkv,
                  const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmOut, const FusedAttnParams p) {
  using namespace fattn;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  float2* ln_part = reinterpret_cast<float2*>(sptr + OFF_MISC);
  float* s_gamma = reinterpret_cast<float*>(sptr + OFF_MISC + 128 * 2 * 8);
  float* s_beta = s_gamma + D;
  float* s_bqkv = s_beta + D;
  float* s_bo = s_bqkv + 3 * D;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M + TM - 1) / TM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWqkv); tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmOut);
    mbar_init(bar(B_XFULL), 1); mbar_init(bar(B_XEMPTY), 1); mbar_init(bar(B_XNREADY), 8);
    mbar_init(bar(B_QKVFULL), 1); mbar_init(bar(B_QKVEMPTY), 8); mbar_init(bar(B_QKREADY), 8);
    mbar_init(bar(B_SFULL), 1); mbar_init(bar(B_PREADY), 8);
    mbar_init(bar(B_OFULL), 1); mbar_init(bar(B_OEMPTY), 8);
    mbar_init(bar(B_PFULL), 1); mbar_init(bar(B_PEMPTY), 8); mbar_init(bar(B_RESFULL), 1); mbar_init(bar(B_RESGO), 1);
    for (int i = 0; i < NSTW; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(bar(B_TMEMSLOT));
  for (int i = threadIdx.x; i < D; i += THREADS) { s_gamma[i] = p.gamma[i]; s_beta[i] = p.beta[i]; s_bo[i] = p.bo[i]; }
  for (int i = threadIdx.x; i < 3 * D; i += THREADS) s_bqkv[i] = p.bqkv[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMSLOT)));

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int ws = 0; uint32_t wph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(bar(B_XEMPTY), (it & 1) ^ 1);
        mbar_arrive_expect_tx(bar(B_XFULL), XN_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_XN + kb * KBLK, &tmX, kb * 64, tile * TM, bar(B_XFULL));
        for (int h = 0; h < NH; ++h)
          for (int kb = 0; kb < KB_X; ++kb) {                // [Wq_h; Wk_h; Wv_h] K-block: 3 boxes of 64 rows
            mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
            mbar_arrive_expect_tx(bar(B_WFULL + ws), STAGE_BYTES);
            const uint32_t dst = sbase + OFF_W + ws * STAGE_BYTES;
            for (int part = 0; part < 3; ++part) tma_load_2d(dst + part * 8192, &tmWqkv, kb * 64, part * D + h * DH, bar(B_WFULL + ws));
            if (++ws == NSTW) { ws = 0; wph ^= 1; }
          }
        for (int kb = 0; kb < KB_X; ++kb) {                  // Wo K-block [192 x 64]
          mbar_wait(bar(B_WEMPTY + ws), wph ^ 1);
          mbar_arrive_expect_tx(bar(B_WFULL + ws), STAGE_BYTES);
          tma_load_2d(sbase + OFF_W + ws * STAGE_BYTES, &tmWo, kb * 64, 0, bar(B_WFULL + ws));
          if (++ws == NSTW) { ws = 0; wph ^= 1; }
        }
        // residual: once the last head's P.V has retired (RESGO) the Q|K|V buffers are dead -> reload the raw x tile there
        mbar_wait(bar(B_RESGO), it & 1);
        mbar_arrive_expect_tx(bar(B_RESFULL), XN_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_QK + kb * KBLK, &tmX, kb * 64, tile * TM, bar(B_RESFULL));
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t id_qkv = umma_idesc_bf16(TM, 192, false, false);
      constexpr uint32_t id_s = umma_idesc_bf16(TM, 128, false, false);
      constexpr uint32_t id_pv = umma_idesc_bf16(TM, 64, false, true);     // B = V_h, MN-major
      int ws = 0; uint32_t wph = 0;
      uint32_t n_head = 0;                       // heads processed so far (barrier phase = n & 1)
      int it = 0;
      auto wgemm = [&](uint32_t a_base, uint32_t d_col) {   // [128 x 192] . W-stage^T over 3 K-blocks, N = 192
        for (int kb = 0; kb < KB_X; ++kb) {
          mbar_wait(bar(B_WFULL + ws), wph);
          tc_fence_after();
          const uint32_t sa = a_base + kb * KBLK, sb = sbase + OFF_W + ws * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + d_col, umma_desc_sw128(sa + k * 32, 16, 1024), umma_desc_sw128(sb + k * 32, 16, 1024), id_qkv,
                      (kb | k) ? 1u : 0u);
          umma_commit(bar(B_WEMPTY + ws));
          if (++ws == NSTW) { ws = 0; wph ^= 1; }
        }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(bar(B_XNREADY), it & 1);
        mbar_wait(bar(B_PEMPTY), (it & 1) ^ 1);            // previous tile's final epilogue has drained TMEM cols 0..191
        tc_fence_after();
        FATT_STAMP(99);
        wgemm(sbase + OFF_XN, COL_QKV);                    // QKV(0)
        umma_commit(bar(B_QKVFULL));
        FATT_STAMP(100);
        for (int h = 0; h < NH; ++h, ++n_head) {
          const uint32_t ph = n_head & 1;
          // ---- S = Q K^T ----
          mbar_wait(bar(B_QKREADY), ph);
          tc_fence_after();
          FATT_STAMP(101 + 8 * h);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + COL_S, umma_desc_sw128(sbase + OFF_QK + k * 32, 16, 1024),
                      umma_desc_sw128(sbase + OFF_QK + KBLK + k * 32, 16, 1024), id_s, k ? 1u : 0u);
          umma_commit(bar(B_SFULL));
          FATT_STAMP(102 + 8 * h);
          // ---- next head's projections overlap the softmax ----
          if (h + 1 < NH) {
            mbar_wait(bar(B_QKVEMPTY), ph);                // accumulator of head h drained (arrived before QKREADY(h))
            tc_fence_after();
            wgemm(sbase + OFF_XN, COL_QKV);
            umma_commit(bar(B_QKVFULL));
            if (h + 2 == NH) umma_commit(bar(B_XEMPTY));   // last read of LN(x): the x buffer frees when it retires
          }
          FATT_STAMP(103 + 8 * h);
          // ---- O = P V ----
          mbar_wait(bar(B_PREADY), ph);
          FATT_STAMP(104 + 8 * h);
          if (n_head > 0) mbar_wait(bar(B_OEMPTY), (n_head - 1) & 1);   // previous O drained from TMEM
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)                       // K = 128 keys: P K-block k/4, 16-key step k%4; V rows 16k..16k+15
            umma_bf16(tmem_base + COL_O, umma_desc_sw128(sbase + OFF_QK + (k >> 2) * KBLK + (k & 3) * 32, 16, 1024),
                      umma_desc_sw128(sbase + OFF_V + k * 2048, 8192, 1024), id_pv, k ? 1u : 0u);
          umma_commit(bar(B_OFULL));
          if (h == NH - 1) umma_commit(bar(B_RESGO));        // Q|K|V buffers dead: the producer may reload the residual tile
          FATT_STAMP(105 + 8 * h);
        }
        // ---- output projection over concat(O_h) ----
        mbar_wait(bar(B_OEMPTY), (n_head - 1) & 1);        // OC complete
        mbar_wait(bar(B_QKVEMPTY), (n_head - 1) & 1);      // cols 0..191 drained by the last head's epilogue
        tc_fence_after();
        FATT_STAMP(130);
        wgemm(sbase + OFF_OC, COL_QKV);
        umma_commit(bar(B_PFULL));
        FATT_STAMP(131);
      }
    }
  } else {
    // =============================== compute warps (2..9) ===============================
    const int cw = warp - 2;
    const int quad = warp & 3;
    const int hf = cw >> 2;
    const int row = quad * 32 + lane;
    const int img = row >> 6;                               // which of the tile's two images this row belongs to
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);
    uint32_t n_head = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const bool stamp = (warp == 2 && lane == 0);
      if (stamp) { bulk_wait_read0(); FATT_STAMP(0); }        // previous tile's TMA store has finished reading the staging smem
      // ---- LayerNorm in place (2 threads per row) ----
      mbar_wait(bar(B_XFULL), it & 1);
      if (stamp) FATT_STAMP(1);
      uint4 v[12];
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int col = hf * 96 + i * 8, kb = col >> 6, ch = (col & 63) >> 3;
        v[i] = *reinterpret_cast<const uint4*>(sptr + OFF_XN + kb * KBLK + row * 128 + ((ch ^ sw) << 4));
        const uint32_t* w = &v[i].x;
#pragma unroll
        for (int j = 0; j < 4; ++j) { float a = bf16_lo(w[j]), b = bf16_hi(w[j]); s += a + b; q += a * a + b * b; }
      }
      ln_part[row * 2 + hf] = make_float2(s, q);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 o = ln_part[row * 2 + (hf ^ 1)];
      const float mean = (s + o.x) * (1.0f / D);
      const float var = fmaxf((q + o.y) * (1.0f / D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const int col = hf * 96 + i * 8, kb = col >> 6, ch = (col & 63) >> 3;
        uint32_t* w = &v[i].x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = col + 2 * j;
          w[j] = pack_bf16((bf16_lo(w[j]) - mean) * rstd * s_gamma[cc] + s_beta[cc], (bf16_hi(w[j]) - mean) * rstd * s_gamma[cc + 1] + s_beta[cc + 1]);
        }
        *reinterpret_cast<uint4*>(sptr + OFF_XN + kb * KBLK + row * 128 + ((ch ^ sw) << 4)) = v[i];
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_XNREADY));
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (stamp) FATT_STAMP(2);

      for (int h = 0; h < NH; ++h, ++n_head) {
        const uint32_t ph = n_head & 1;
        // ---- QKV epilogue: thread (row, hf) owns accumulator columns [hf*96, hf*96+96) of [Q | K | V] ----
        mbar_wait(bar(B_QKVFULL), ph);
        tc_fence_after();
        if (stamp) FATT_STAMP(10 + 8 * h);
#pragma unroll 1
        for (int cc = 0; cc < 3; ++cc) {
          const int col = hf * 96 + cc * 32;               // 0..191, 32-wide pieces never straddle Q/K/V (64-wide)
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + tm_lane + COL_QKV + col, r);
          tmem_ld_wait();
          const int part = col >> 6, c0 = col & 63;        // part: 0 Q, 1 K, 2 V
          const float* bias = s_bqkv + part * D + h * DH + c0;
          uint8_t* dst = sptr + (part == 2 ? OFF_V : OFF_QK + part * KBLK) + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = 8 * j + 2 * k;
              w[k] = pack_bf16(__uint_as_float(r[e]) + bias[e], __uint_as_float(r[e + 1]) + bias[e + 1]);
            }
            *reinterpret_cast<uint4*>(dst + ((((c0 >> 3) + j) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar(B_QKVEMPTY)); mbar_arrive(bar(B_QKREADY)); }
        if (stamp) FATT_STAMP(11 + 8 * h);

        // ---- softmax over this row's 64 keys (its own image) ----
        mbar_wait(bar(B_SFULL), ph);
        tc_fence_after();
        if (stamp) FATT_STAMP(12 + 8 * h);
        // thread (row, hf) owns 32 of the row's 64 keys; row max / sum are exchanged with the partner thread through smem
        uint32_t sv[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_S + img * 64 + hf * 32, sv);
        tmem_ld_wait();
        const float sl2 = 0.125f * 1.4426950408889634f;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]));
        ln_part[row * 2 + hf].x = mx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(mx, ln_part[row * 2 + (hf ^ 1)].x);
        const float moff = mx * sl2;
        float sum = 0.f;
        float e_own[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { e_own[j] = exp2f(__uint_as_float(sv[j]) * sl2 - moff); sum += e_own[j]; }
        ln_part[row * 2 + hf].y = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = 1.0f / (sum + ln_part[row * 2 + (hf ^ 1)].y);
        // P: K-block `img` holds this image's keys; this thread writes keys [hf*32, hf*32+32) of its row and zeroes the
        // same span of the other image's K-block (P aliases the Q/K buffers, which S has finished reading: SFULL)
        uint8_t* prow = sptr + OFF_QK + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t chunk = (uint32_t)(hf * 4 + j);
          *reinterpret_cast<uint4*>(prow + img * KBLK + ((chunk ^ sw) << 4)) =
              make_uint4(pack_bf16(e_own[8 * j] * inv, e_own[8 * j + 1] * inv), pack_bf16(e_own[8 * j + 2] * inv, e_own[8 * j + 3] * inv),
                         pack_bf16(e_own[8 * j + 4] * inv, e_own[8 * j + 5] * inv), pack_bf16(e_own[8 * j + 6] * inv, e_own[8 * j + 7] * inv));
          *reinterpret_cast<uint4*>(prow + (img ^ 1) * KBLK + ((chunk ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PREADY));
        if (stamp) FATT_STAMP(13 + 8 * h);

        // ---- O_h epilogue -> K-block h of the projection's A operand ----
        mbar_wait(bar(B_OFULL), ph);
        tc_fence_after();
        if (stamp) FATT_STAMP(14 + 8 * h);
        uint32_t ro[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_O + hf * 32, ro);
        tmem_ld_wait();
        uint8_t* orow = sptr + OFF_OC + h * KBLK + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(orow + (((uint32_t)(hf * 4 + j) ^ sw) << 4)) =
              make_uint4(pack_bf16(__uint_as_float(ro[8 * j]), __uint_as_float(ro[8 * j + 1])), pack_bf16(__uint_as_float(ro[8 * j + 2]), __uint_as_float(ro[8 * j + 3])),
                         pack_bf16(__uint_as_float(ro[8 * j + 4]), __uint_as_float(ro[8 * j + 5])), pack_bf16(__uint_as_float(ro[8 * j + 6]), __uint_as_float(ro[8 * j + 7])));
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_OEMPTY));
        if (stamp) FATT_STAMP(15 + 8 * h);
      }

      // ---- final epilogue: out = proj + bo + x  (residual tile in smem, result staged in smem, TMA store) ----
      mbar_wait(bar(B_PFULL), it & 1);
      mbar_wait(bar(B_RESFULL), it & 1);
      tc_fence_after();
      if (stamp) FATT_STAMP(60);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int col = hf * 96 + cc * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_QKV + col, r);
        tmem_ld_wait();
        const int kb = col >> 6, ch0 = (col & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = (uint32_t)(kb * KBLK + row * 128) + ((((uint32_t)(ch0 + j)) ^ sw) << 4);
          const uint4 xr = *reinterpret_cast<const uint4*>(sptr + OFF_QK + off);
          const uint32_t* xw = &xr.x;
          uint32_t ow[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = 8 * j + 2 * k;
            ow[k] = pack_bf16(__uint_as_float(r[e]) + s_bo[col + e] + bf16_lo(xw[k]), __uint_as_float(r[e + 1]) + s_bo[col + e + 1] + bf16_hi(xw[k]));
          }
          *reinterpret_cast<uint4*>(sptr + OFF_OC + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (stamp) {
        for (int kb = 0; kb < KB_X; ++kb) tma_store_2d(&tmOut, sbase + OFF_OC + kb * KBLK, kb * 64, tile * TM);
        bulk_commit();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PEMPTY));
      if (stamp) FATT_STAMP(61);
    }
  }
  if (warp == 2 && lane == 0) bulk_wait0();                  // outstanding TMA stores complete before the CTA exits
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<fattn::TMEM_COLS>(tmem_base);
  }
}

static long long* g_fattn_dbg = nullptr;
void fused_attn_set_debug(long long* buf) { g_fattn_dbg = buf; }

bool fused_attn_supported(int D, int heads, int tokens) { return D == fattn::D && heads == fattn::NH && tokens == 64; }

int launch_fused_attn(cudaStream_t stream, const __nv_bfloat16* x, __nv_bfloat16* out, const float* gamma, const float* beta,
                      const __nv_bfloat16* wqkv, const float* bqkv, const __nv_bfloat16* wo, const float* bo, int M, int D, int heads, float eps) {
  using namespace fattn;
  if (D != fattn::D || heads != NH) { set_last_error("fused_attn: only D=192, 3 heads"); return VITMARL_EINVAL; }
  if (M <= 0) return VITMARL_OK;
  if (M % 64) { set_last_error("fused_attn: tokens must be whole images of 64"); return VITMARL_EINVAL; }
  CUtensorMap tmX, tmWqkv, tmWo, tmOut;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmX, x, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmOut, out, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmWqkv, wqkv, 3 * D, D, (uint64_t)D * 2, 64, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmWo, wo, D, D, (uint64_t)D * 2, D, 64))) return rc;
  FusedAttnParams p{M, x, out, gamma, beta, bqkv, bo, eps, g_fattn_dbg};
  cudaError_t e = cudaFuncSetAttribute(fused_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e);
  const int tiles = (M + TM - 1) / TM;
  fused_attn_kernel<<<min(tiles, num_sms()), THREADS, SMEM_BYTES, stream>>>(tmX, tmWqkv, tmWo, tmOut, p);
  return check_cuda(cudaGetLastError());
}

}  // namespace vitmarl
