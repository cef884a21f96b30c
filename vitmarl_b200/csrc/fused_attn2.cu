// Fused attention block for D = 192, 3 heads x 64, T = 64 tokens (ViT-Tiny), inference path -- second generation:
//     out = x + Wo . concat_h( softmax(Q_h K_h^T / 8) V_h ) + bo,    [Q|K|V] = LayerNorm(x) . Wqkv^T + bqkv
// Same maths as fused_attn.cu (one kernel per layer, a 128-token tile = two images per CTA), restructured around what
// bounded that kernel (8 compute warps marching in lock step through LN -> epilogue -> softmax -> epilogue, a single
// thread paying ~16 uniform-datapath instructions per tcgen05.mma, every operand re-read from shared memory):
//   * WARP SPECIALISATION: 8 softmax warps (two threads per query row, 32 keys each; row max / sum exchanged through smem,
//     1 / sum left in smem for the O epilogue), 12 conversion warps (thread = (row, part): part 0/1/2 owns Q/K/V of the QKV
//     epilogue and one 64-column K-block of the LayerNorm / final epilogue; parts 0/1 also run the O epilogue) and TWO MMA
//     issuing warps (projections / attention core) run concurrently, so the softmax of head h overlaps the QKV epilogue of
//     head h+1 and the next tile's LayerNorm overlaps this tile's projection;
//   * OPERANDS IN TENSOR MEMORY: LayerNorm writes its bf16 result straight into TMEM and the QKV projections read it as
//     the A operand (tcgen05.mma [d], [a_tmem], b_desc); softmax writes P over the S accumulator in place and P.V reads
//     it from TMEM; Q_h goes to TMEM as well (A operand of S).  Shared memory holds only what must be a B operand
//     (weights, K, V) plus concat(O);
//   * FOLDED PARAMETERS (vit_fold.cu): LayerNorm scale/shift, the 1/8 . log2(e) softmax scale and every bias except
//     Q's are folded into the weights (the K bias cancels inside the softmax, the V bias moves into the output bias),
//     so the epilogues are convert-and-store;
//   * the x tile is consumed by the LayerNorm only (the residual is re-read from L2 in the final epilogue), so the next
//     tile's TMA load is issued as soon as the LayerNorm has read the current one (reading x for the LayerNorm straight
//     from L2 instead was tried: its latency lands on the tile-boundary critical path, 146 -> 180 us);
//   * tile boundary: the next tile's first QKV projection is issued BEFORE this tile's output projection, which
//     accumulates into the (dead) S|O columns, so the tensor pipe never drains between tiles.
// TMEM columns: XN 0..95 (LN(x), bf16 pairs) | QKV 96..287 | S / P 288..415 | O 416..479 | Q_h 480..511 (bf16 pairs);  projection
// accumulator = 288..479.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace fattn2 {
constexpr int D = 192, NH = 3, DH = 64, TM = 128;
constexpr int KB_X = D / 64;                       // 3 K-blocks
constexpr int NSTW = 3;                            // weight ring stages ([192 x 64] K-blocks, 24 KB)
constexpr int STAGE_BYTES = 192 * 128;             // 24 KB
constexpr int KBLK = TM * 128;                     // one [128 x 64] bf16 K-block = 16 KB
constexpr int X_BYTES = KB_X * KBLK;               // 48 KB
constexpr int OFF_X = 0;                           // raw x tile (LayerNorm input)
constexpr int OFF_K = OFF_X + X_BYTES;             // (Q_h lives in tensor memory)
constexpr int OFF_V = OFF_K + KBLK;                // V_h [128 keys x 64 d]
constexpr int OFF_OC = OFF_V + KBLK;               // concat(O_h) [128 x 192] = 3 K-blocks; reused as the output staging tile
constexpr int OFF_W = OFF_OC + 3 * KBLK;
constexpr int OFF_BAR = OFF_W + NSTW * STAGE_BYTES;
constexpr int OFF_MISC = OFF_BAR + 256;
constexpr int MISC_BYTES = 128 * 3 * 8 + D * 2 + D * 4 + 128 * 2 * 8 + 2 * 128 * 4;   // LN partials [128][3] float2, bq' (bf16), bo' (fp32), softmax partials [128][2] float2, 1/sum [2][128]
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int W_TMA = 0, W_MMA = 1, W_IO = 2, W_MMA2 = 3, W_SM0 = 4, W_CV0 = 12;   // + 8 softmax warps 4..11, 12 conversion warps 12..23
constexpr int N_CV = 12, N_SM = 8, N_OE = 4;   // conversion / softmax warps; conversion warps that run the O epilogue (part 2: the V warps,
                                                // which wait for P.V anyway -- the Q / K warps stay decoupled from the softmax chain)
constexpr int THREADS = 32 * (W_CV0 + N_CV);       // 768
constexpr int TMEM_COLS = 512;
constexpr int COL_XN = 0, COL_QKV = 96, COL_S = 288, COL_O = 416, COL_PROJ = 288, COL_Q = 480;
enum { B_XFULL = 0, B_XFREE, B_XNREADY, B_XNFREE, B_QKVFULL, B_QKVEMPTY, B_QKREADY, B_VREADY, B_SFULL, B_PREADY, B_OFULL, B_OCREADY,
       B_PROJFULL, B_PROJEMPTY, B_OUTREADY, B_OCFREE, B_OCDONE, B_SISSUED, B_WFULL, B_WEMPTY = B_WFULL + NSTW, B_TMEMSLOT = B_WEMPTY + NSTW, B_COUNT };
static_assert(B_COUNT * 8 <= 256, "barrier area");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace fattn2

struct FusedAttn2Params {
  int M;
  const __nv_bfloat16* x;       // [M, D] residual stream in (also re-read from L2 for the residual add)
  const uint32_t* bqp;          // [D/2] folded Q bias (scaled), packed bf16x2
  const float* bo;              // [D] folded output bias (bo + Wo . bv')
  float eps;
  long long* dbg;
  int inplace;                  // out aliases x: the residual add is done by a TMA reduce-add store (x += delta), no residual load
  int flags;                    // tuning switches: bit 0 projection of head n+1 held back until S(n) has retired; bit 1 output projection
                                // (and its weights) before the next tile's first QKV projection
};

#define FA2_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && j == 1 && lane == 0) p.dbg[(slot)] = clock64(); } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(fattn2::THREADS, 1)
fused_attn2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWqkv,
                   const __grid_constant__ CUtensorMap tmWo, const __grid_constant__ CUtensorMap tmOut, const FusedAttn2Params p) {
  using namespace fattn2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  float2* ln_part = reinterpret_cast<float2*>(sptr + OFF_MISC);                     // [128][3]
  uint32_t* s_bqp = reinterpret_cast<uint32_t*>(sptr + OFF_MISC + 128 * 3 * 8);     // [96] packed bf16x2
  float* s_bo = reinterpret_cast<float*>(s_bqp + D / 2);
  float2* sm_part = reinterpret_cast<float2*>(s_bo + D);                            // [128][2] (max, sum) partials
  float* s_inv = reinterpret_cast<float*>(sm_part + 128 * 2);                       // [2][128] 1 / sum per row, double buffered by head parity

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M + TM - 1) / TM;
  const int nt = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_row = [&](int j) { return ((int)blockIdx.x + j * (int)gridDim.x) * TM; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWqkv); tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmOut);
    mbar_init(bar(B_XFULL), 1); mbar_init(bar(B_XFREE), N_CV); mbar_init(bar(B_XNREADY), N_CV); mbar_init(bar(B_XNFREE), 1);
    mbar_init(bar(B_QKVFULL), 1); mbar_init(bar(B_QKVEMPTY), N_CV); mbar_init(bar(B_QKREADY), 8); mbar_init(bar(B_VREADY), 4);
    mbar_init(bar(B_SFULL), 1); mbar_init(bar(B_PREADY), N_SM); mbar_init(bar(B_OFULL), 1); mbar_init(bar(B_OCREADY), N_OE);
    mbar_init(bar(B_PROJFULL), 1); mbar_init(bar(B_PROJEMPTY), N_SM); mbar_init(bar(B_OUTREADY), N_SM); mbar_init(bar(B_OCFREE), 1); mbar_init(bar(B_OCDONE), N_OE); mbar_init(bar(B_SISSUED), 1);
    for (int i = 0; i < NSTW; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
    fence_mbar_init();
  }
  if (warp == W_MMA) tmem_alloc<TMEM_COLS>(bar(B_TMEMSLOT));
  for (int i = threadIdx.x; i < D / 2; i += THREADS) s_bqp[i] = p.bqp[i];
  for (int i = threadIdx.x; i < D / 2; i += THREADS) s_bo[i] = __uint_as_float(pack_bf16(p.bo[2 * i], p.bo[2 * i + 1]));   // packed bf16x2
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMSLOT)));
  // PDL: everything above (and the weight ring of W_TMA, which only reads parameters) overlaps the tail of the
  // previous kernel; the residual stream is touched only after it has completed.
  if (warp != W_TMA) { griddep_wait(); griddep_launch_dependents(); }

  if (warp == W_TMA) {
    // =============================== weight ring (in MMA issue order) ===============================
    if (lane == 0 && nt > 0) {
      int ws = 0; uint32_t wph = 0;
      auto load_qkv = [&](int h) {                           // [Wq_h; Wk_h; Wv_h] K-blocks: 3 boxes of 64 rows each
        for (int kb = 0; kb < KB_X; ++kb) {
          mbar_wait_guard(bar(B_WEMPTY + ws), wph ^ 1);
          mbar_arrive_expect_tx(bar(B_WFULL + ws), STAGE_BYTES);
          const uint32_t dst = sbase + OFF_W + ws * STAGE_BYTES;
          for (int part = 0; part < 3; ++part) tma_load_2d(dst + part * 8192, &tmWqkv, kb * 64, part * D + h * DH, bar(B_WFULL + ws));
          if (++ws == NSTW) { ws = 0; wph ^= 1; }
        }
      };
      load_qkv(0);
      for (int j = 0; j < nt; ++j) {
        load_qkv(1);
        load_qkv(2);
        const bool wo_first = (p.flags & 2) != 0;
        if (!wo_first && j + 1 < nt) load_qkv(0);
        for (int kb = 0; kb < KB_X; ++kb) {                  // Wo K-block [192 x 64]
          mbar_wait_guard(bar(B_WEMPTY + ws), wph ^ 1);
          mbar_arrive_expect_tx(bar(B_WFULL + ws), STAGE_BYTES);
          tma_load_2d(sbase + OFF_W + ws * STAGE_BYTES, &tmWo, kb * 64, 0, bar(B_WFULL + ws));
          if (++ws == NSTW) { ws = 0; wph ^= 1; }
        }
        if (wo_first && j + 1 < nt) load_qkv(0);
      }
    }
  } else if (warp == W_IO) {
    // =============================== x tiles in, output tiles out ===============================
    if (lane == 0 && nt > 0) {
      auto load_x = [&](int j) {
        mbar_arrive_expect_tx(bar(B_XFULL), X_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_X + kb * KBLK, &tmX, kb * 64, tile_row(j), bar(B_XFULL));
      };
      load_x(0);
      if (nt > 1) { mbar_wait_guard(bar(B_XFREE), 0); load_x(1); }
      for (int j = 0; j < nt; ++j) {
        // the x tile of tile j+2 is requested as soon as LayerNorm(j+1) has read the buffer (in tile j's tail) and BEFORE this
        // tile's output store is issued: a load queued behind the 48 KB reduce-add store reached the LayerNorm ~3000 cycles late
        if (j + 2 < nt) { mbar_wait_guard(bar(B_XFREE), (j + 1) & 1); load_x(j + 2); }
        mbar_wait_guard(bar(B_OUTREADY), j & 1);
        for (int kb = 0; kb < KB_X; ++kb) {
          if (p.inplace) tma_reduce_add_2d(&tmOut, sbase + OFF_OC + kb * KBLK, kb * 64, tile_row(j));
          else tma_store_2d(&tmOut, sbase + OFF_OC + kb * KBLK, kb * 64, tile_row(j));
        }
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar(B_OCFREE));                                                   // the staging tile may be overwritten
      }
      bulk_wait0();
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer A: the weight-streaming projections (QKV_h, output projection) ===============================
    // Two issuing warps feed the one tensor pipe so that a projection waiting for its accumulator / weights never holds
    // back an S or P.V product that is ready (and vice versa); every cross-warp hazard is covered by an mbarrier.
    if (nt > 0) {
      constexpr uint32_t id_n192 = umma_idesc_bf16(TM, 192, false, false);
      int ws = 0; uint32_t wph = 0;
      auto qkv_gemm = [&]() {                                // QKV acc = XN[tmem] . W-stages^T, N = 192, A in tensor memory
        for (int kb = 0; kb < KB_X; ++kb) {
          mbar_wait_guard(bar(B_WFULL + ws), wph);
          tc_fence_after();
          const uint32_t lb = umma_desc_lo(sbase + OFF_W + ws * STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tmem_base + COL_QKV, tmem_base + COL_XN + (kb * 4 + k) * 8, umma_desc_from_lo(lb + 2 * k), id_n192, (kb | k) ? 1u : 0u);
            umma_commit(bar(B_WEMPTY + ws));
          }
          __syncwarp();
          if (++ws == NSTW) { ws = 0; wph ^= 1; }
        }
      };
      mbar_wait_guard(bar(B_XNREADY), 0);
      tc_fence_after();
      qkv_gemm();
      if (elect_one()) umma_commit(bar(B_QKVFULL));
      __syncwarp();
      uint32_t n = 0;                                          // global head counter
      for (int j = 0; j < nt; ++j) {
        for (int h = 0; h + 1 < NH; ++h, ++n) {
          mbar_wait_guard(bar(B_QKVEMPTY), n & 1);             // accumulator of head n drained (early in its epilogue)
          // default: wait until S(n) is in the pipe so that the long projection queues BEHIND the short S product;
          // flag bit 2: issue the projection at once (the phase is then consumed after the issue -- every phase must be)
          if (!(p.flags & 4)) mbar_wait_guard(bar(B_SISSUED), n & 1);
          if (p.flags & 1) mbar_wait_guard(bar(B_SFULL), n & 1); // S(n) first: the projection then overlaps the softmax, not S / P.V
          tc_fence_after();
          FA2_STAMP(101 + 4 * h);
          qkv_gemm();
          if (elect_one()) {
            umma_commit(bar(B_QKVFULL));
            if (h + 2 == NH) umma_commit(bar(B_XNFREE));       // last read of LN(x): the next tile's LayerNorm may overwrite XN
          }
          __syncwarp();
          if (p.flags & 4) mbar_wait_guard(bar(B_SISSUED), n & 1);
        }
        ++n;                                                   // n = 3 (j + 1): heads of this tile all issued
        // ---- tile boundary: this tile's output projection and the next tile's first QKV projection (order = ring order) ----
        const bool wo_first = (p.flags & 2) != 0;
        auto next_qkv = [&]() {
          mbar_wait_guard(bar(B_SISSUED), (n - 1) & 1);        // (every phase is consumed, also on the last tile)
          if (j + 1 < nt) {
            mbar_wait_guard(bar(B_XNREADY), (j + 1) & 1);
            mbar_wait_guard(bar(B_QKVEMPTY), (n - 1) & 1);
            tc_fence_after();
            qkv_gemm();
            if (elect_one()) umma_commit(bar(B_QKVFULL));
            __syncwarp();
          }
        };
        if (!wo_first) next_qkv();
        // concat(O) complete (implies P.V of the last head has retired).  A per-TILE barrier: this warp does not follow the
        // per-head OCREADY phases, and a parity wait on a barrier that may be two or more phases ahead aliases.
        mbar_wait_guard(bar(B_OCDONE), j & 1);
        tc_fence_after();
        FA2_STAMP(130);
        for (int kb = 0; kb < KB_X; ++kb) {
          mbar_wait_guard(bar(B_WFULL + ws), wph);
          tc_fence_after();
          const uint32_t la = umma_desc_lo(sbase + OFF_OC + kb * KBLK), lb = umma_desc_lo(sbase + OFF_W + ws * STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + COL_PROJ, umma_desc_from_lo(la + 2 * k), umma_desc_from_lo(lb + 2 * k), id_n192, (kb | k) ? 1u : 0u);
            umma_commit(bar(B_WEMPTY + ws));
          }
          __syncwarp();
          if (++ws == NSTW) { ws = 0; wph ^= 1; }
        }
        if (elect_one()) umma_commit(bar(B_PROJFULL));
        __syncwarp();
        FA2_STAMP(131);
        if (wo_first) next_qkv();
      }
    }
  } else if (warp == W_MMA2) {
    // =============================== MMA issuer B: the attention core (S = Q K^T, O = P V) ===============================
    if (nt > 0) {
      constexpr uint32_t id_s = umma_idesc_bf16(TM, 128, false, false);
      constexpr uint32_t id_pv = umma_idesc_bf16(TM, 64, false, true);      // B = V_h, MN-major
      const uint32_t lb_k = umma_desc_lo(sbase + OFF_K), lv = umma_desc_lo(sbase + OFF_V, 8192);
      auto issue_s = [&](uint32_t m) {        // elected lane only: S = Q K^T (Q in tensor memory; both images, the diagonal 64x64 blocks are used)
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base + COL_S, tmem_base + COL_Q + 8 * k, umma_desc_from_lo(lb_k + 2 * k), id_s, k ? 1u : 0u);
        umma_commit(bar(B_SFULL));
        mbar_arrive(bar(B_SISSUED));
      };
      uint32_t n = 0;
      bool s_issued = false;                   // S(n) already went out back to back with P.V(n-1)
      for (int j = 0; j < nt; ++j) {
        for (int h = 0; h < NH; ++h, ++n) {
          const uint32_t ph = n & 1;
          // ---- S(n).  Issued after P.V(n-1) by the same thread, so the in-order tensor pipe has finished reading P before S
          //      overwrites it ----
          if (!s_issued) {
            if (h == 0) mbar_wait_guard(bar(B_PROJEMPTY), (j & 1) ^ 1);    // previous tile's final epilogue drained cols 288..479
            mbar_wait_guard(bar(B_QKREADY), ph);
            tc_fence_after();
            FA2_STAMP(100 + 4 * h);
            if (p.dbg && blockIdx.x == 0 && h == 0 && j < 16 && lane == 0) p.dbg[140 + j] = clock64();   // per-tile period
            if (elect_one()) issue_s(n);
            __syncwarp();
          }
          // ---- O = P V  (P in tensor memory over the S columns, V an MN-major B operand).  The barriers that complete early
          //      first: only one wait latency is left once the softmax arrives on PREADY; if the next head's Q / K are already
          //      in place, S(n+1) goes out in the same breath ----
          mbar_wait_guard(bar(B_VREADY), ph);
          if (n > 0) mbar_wait_guard(bar(B_OCREADY), (n - 1) & 1);          // previous O drained from TMEM
          // (warp-uniform: the phase is complete as soon as any lane has observed it)
          const bool next_ready = (h + 1 < NH) && __any_sync(0xffffffffu, mbar_try_wait(bar(B_QKREADY), ph ^ 1));
          mbar_wait_guard(bar(B_PREADY), ph);
          tc_fence_after();
          FA2_STAMP(102 + 4 * h);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)                                      // 16 keys per step: P columns 8k.., V rows 16k.. (2048 B)
              umma_bf16_ts(tmem_base + COL_O, tmem_base + COL_S + 8 * k, umma_desc_from_lo(lv + k * (2048 >> 4)), id_pv, k ? 1u : 0u);
            umma_commit(bar(B_OFULL));
            if (next_ready) issue_s(n + 1);
          }
          __syncwarp();
          s_issued = next_ready;
        }
      }
    }
  } else if (warp < W_CV0) {
    // =============================== softmax warps (4..11): thread = (query row, half of its 64 keys) ===============================
    const int quad = warp & 3;
    const int half = (warp - W_SM0) >> 2;
    const int row = quad * 32 + lane;
    const int img = row >> 6;
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    uint32_t n = 0;
    for (int j = 0; j < nt; ++j) {
      for (int h = 0; h < NH; ++h, ++n) {
        const uint32_t ph = n & 1;
        mbar_wait_guard(bar(B_SFULL), ph);
        tc_fence_after();
        if (warp == W_SM0) FA2_STAMP(10 + 4 * h);
        uint32_t sv[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_S + img * 64 + half * 32, sv);
        tmem_ld_wait();
        // scores already carry 1/8 . log2(e) (folded into Wq / bq).  Each thread exponentiates against the max of ITS 32 keys;
        // the pair then exchanges (max, sum) once and rescales -- one barrier per head instead of two (the barrier also
        // orders every S read of the pair before the in-place P writes)
        float mx = __uint_as_float(sv[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = ex2_approx(__uint_as_float(sv[i]) - mx);
          sum += e;
          sv[i] = __float_as_uint(e);
        }
        sm_part[row * 2 + half] = make_float2(mx, sum);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float2 other = sm_part[row * 2 + (half ^ 1)];
        const float m = fmaxf(mx, other.x);
        const float f = ex2_approx(mx - m);
        if (half == 0) s_inv[ph * 128 + row] = 1.0f / (sum * f + other.y * ex2_approx(other.x - m));   // read by the O epilogue after OFULL
        uint32_t pw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pw[i] = pack_bf16(__uint_as_float(sv[2 * i]) * f, __uint_as_float(sv[2 * i + 1]) * f);
        // P (unnormalised, bf16 pairs) over this row's S columns: keys of its own image, zeros for the other image's keys
        tmem_st_32x16(tmem_base + tm_lane + COL_S + img * 32 + half * 16, pw);
        tmem_st_32x16_fill(tmem_base + tm_lane + COL_S + (img ^ 1) * 32 + half * 16, 0u);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_PREADY));
        if (warp == W_SM0) FA2_STAMP(11 + 4 * h);
      }
      // ---- final epilogue of tile j: out = proj + bo' (+ x when not in place), thread = (row, 96 of the 192 columns), staged over
      //      concat(O) (dead once PROJFULL has fired), stored / reduce-added by the IO warp ----
      {
        const int col0 = half * 96;
        const uint32_t sw = (uint32_t)(row & 7);
        const int grow = tile_row(j) + row;
        mbar_wait_guard(bar(B_PROJFULL), j & 1);
        tc_fence_after();
        if (warp == W_SM0) FA2_STAMP(60);
        const uint4* gx = reinterpret_cast<const uint4*>(p.x + (size_t)grow * D + col0);
#pragma unroll
        for (int g3 = 0; g3 < 3; ++g3) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + tm_lane + COL_PROJ + col0 + g3 * 32, r);
          tmem_ld_wait();
          if (g3 == 2) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_PROJEMPTY));                      // accumulator drained: the next tile's S may overwrite it
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = col0 + g3 * 32 + 8 * i, kb = col >> 6, c8 = (col & 63) >> 3;
            const uint4 xv = (!p.inplace && grow < p.M) ? gx[g3 * 4 + i] : make_uint4(0u, 0u, 0u, 0u);
            const uint4 bv = *reinterpret_cast<const uint4*>(s_bo + (col >> 1));
            uint4 ov;
            ov.x = bf16x2_add(bf16x2_add(pack_bf16(__uint_as_float(r[8 * i + 0]), __uint_as_float(r[8 * i + 1])), bv.x), xv.x);
            ov.y = bf16x2_add(bf16x2_add(pack_bf16(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])), bv.y), xv.y);
            ov.z = bf16x2_add(bf16x2_add(pack_bf16(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])), bv.z), xv.z);
            ov.w = bf16x2_add(bf16x2_add(pack_bf16(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])), bv.w), xv.w);
            *reinterpret_cast<uint4*>(sptr + OFF_OC + kb * KBLK + row * 128 + ((((uint32_t)c8) ^ sw) << 4)) = ov;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_OUTREADY));
        if (warp == W_SM0) FA2_STAMP(61);
      }
    }
  } else {
    // =============================== conversion warps (12..23): thread = (row, part) ===============================
    const int quad = warp & 3;
    const int part = (warp - W_CV0) >> 2;                    // 0 Q / 1 K / 2 V of the QKV epilogue; K-block of LN / final epilogue
    const int row = quad * 32 + lane;
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);

    // ---- LayerNorm (scale/shift folded into Wqkv'/bq'): x K-block `part` of this row -> bf16 pairs in TMEM (XN) ----
    auto layer_norm = [&](int jn) {
      const int j = jn - 1;                                   // (stamps are keyed on the tile that is current while this runs)
      if (warp == W_CV0) FA2_STAMP(51);
      mbar_wait_guard(bar(B_XFULL), jn & 1);
      if (warp == W_CV0) FA2_STAMP(52);
      const uint8_t* xr = sptr + OFF_X + part * KBLK + row * 128;
      uint4 v[8];
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        v[c] = *reinterpret_cast<const uint4*>(xr + ((((uint32_t)c) ^ sw) << 4));
        stats_bf16x2(v[c].x, s, q); stats_bf16x2(v[c].y, s, q); stats_bf16x2(v[c].z, s, q); stats_bf16x2(v[c].w, s, q);
      }
      ln_part[row * 3 + part] = make_float2(s, q);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_XFREE));                            // x tile consumed (values live in registers)
      asm volatile("bar.sync 1, 384;" ::: "memory");
      if (warp == W_CV0) FA2_STAMP(53);
#pragma unroll
      for (int k = 1; k < 3; ++k) { const float2 o = ln_part[row * 3 + (part + k) % 3]; s += o.x; q += o.y; }
      const float mean = s * (1.0f / D);
      const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
      const uint32_t rstd_b = pack_bf16(rsqrtf(var + p.eps), 0.f);
      const float nmr = -mean * bf16_lo(rstd_b);
      uint32_t w[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        w[4 * c + 0] = norm_bf16x2(v[c].x, rstd_b, nmr); w[4 * c + 1] = norm_bf16x2(v[c].y, rstd_b, nmr);
        w[4 * c + 2] = norm_bf16x2(v[c].z, rstd_b, nmr); w[4 * c + 3] = norm_bf16x2(v[c].w, rstd_b, nmr);
      }
      if (jn > 0) { mbar_wait_guard(bar(B_XNFREE), (jn - 1) & 1); tc_fence_after(); }   // previous tile's last QKV projection has read XN
      if (warp == W_CV0) FA2_STAMP(54);
      tmem_st_32x32(tmem_base + tm_lane + COL_XN + part * 32, w);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_XNREADY));
      asm volatile("bar.sync 1, 384;" ::: "memory");                       // ln_part reuse safety
    };

    // ---- O_h epilogue (part 2 warps, thread = row): normalise by the row's 1 / sum, convert, K-block h of the projection's A operand ----
    auto o_epilogue = [&](int j, int h, uint32_t n) {
      if (part != 2) return;
      mbar_wait_guard(bar(B_OFULL), n & 1);
      if (h == 0) mbar_wait_guard(bar(B_OCFREE), (j & 1) ^ 1);             // previous tile's output store has read the staging tile
      tc_fence_after();
      const float inv = s_inv[(n & 1) * 128 + row];
      uint8_t* orow = sptr + OFF_OC + h * KBLK + row * 128;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_O + hf * 32, r);
        tmem_ld_wait();
        if (hf == 1) tc_fence_before();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(orow + ((((uint32_t)(hf * 4 + c)) ^ sw) << 4)) =
              make_uint4(pack_bf16(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv),
                         pack_bf16(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv),
                         pack_bf16(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv),
                         pack_bf16(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(B_OCREADY)); if (h == NH - 1) mbar_arrive(bar(B_OCDONE)); }
      if (warp == W_CV0 + 8) FA2_STAMP(13 + 4 * h);
    };

    // ---- QKV epilogue of head n: 64 accumulator columns of Q_h, K_h or V_h -> bf16 -> swizzled operand tile in smem ----
    auto qkv_epilogue = [&](int j, int h, uint32_t n) {
      mbar_wait_guard(bar(B_QKVFULL), n & 1);
      tc_fence_after();
      if (warp == W_CV0) FA2_STAMP(30 + 2 * h);
      uint32_t w[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + tm_lane + COL_QKV + part * 64 + hf * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) w[hf * 16 + i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      }
      if (warp == W_CV0 && h == 1) FA2_STAMP(70);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_QKVEMPTY));                         // accumulator drained: the next projection may start
      if (part == 0) {
        const uint4* bq = reinterpret_cast<const uint4*>(s_bqp + h * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 b = bq[c];
          w[4 * c] = bf16x2_add(w[4 * c], b.x); w[4 * c + 1] = bf16x2_add(w[4 * c + 1], b.y);
          w[4 * c + 2] = bf16x2_add(w[4 * c + 2], b.z); w[4 * c + 3] = bf16x2_add(w[4 * c + 3], b.w);
        }
      }
      // operand buffers of the previous head must be dead: Q, K once S has completed; V once P.V has completed
      if (n > 0) {
        if (part == 2) mbar_wait_guard(bar(B_OFULL), (n - 1) & 1);
        else mbar_wait_guard(bar(B_SFULL), (n - 1) & 1);
      }
      if (warp == W_CV0 && h == 1) FA2_STAMP(71);
      if (part == 0) {
        // Q_h goes to TENSOR MEMORY (32 columns of bf16 pairs, A operand of S = Q K^T): shared memory is the busiest resource of
        // this kernel (ncu: LSU + tensor-core wavefronts ~50 % of the data pipe on average, before the TMA writes), and Q through
        // shared memory cost 768 of its ~10 k wavefronts per tile
        tmem_st_32x32(tmem_base + tm_lane + COL_Q, w);
        tmem_st_wait();
        tc_fence_before();
      } else {
        uint8_t* dst = sptr + (part == 1 ? OFF_K : OFF_V) + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(dst + ((((uint32_t)c) ^ sw) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
      }
      if (warp == W_CV0 && h == 1) FA2_STAMP(72);
      if (part != 0) fence_proxy_async_smem();
      if (warp == W_CV0 && h == 1) FA2_STAMP(73);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(part == 2 ? B_VREADY : B_QKREADY));
      if (warp == W_CV0) FA2_STAMP(31 + 2 * h);
    };

    // (A version software-pipelined over tiles -- next tile's first QKV epilogue before this tile's final epilogue, next
    // LayerNorm after head 1 -- measured slower, 146 -> 171 us: it lengthens the conversion warps' serial chain.)
    if (nt > 0) layer_norm(0);
    uint32_t n = 0;
    for (int j = 0; j < nt; ++j) {
      for (int h = 0; h < NH; ++h, ++n) {
        qkv_epilogue(j, h, n);
        if (h > 0) o_epilogue(j, h - 1, n - 1);
      }
      // next tile's LayerNorm (its x tile was prefetched during this tile) overlaps this tile's P.V / projection
      if (j + 1 < nt) layer_norm(j + 1);
      if (warp == W_CV0) FA2_STAMP(50);
      o_epilogue(j, NH - 1, n - 1);
      // (the final epilogue of the tile runs on the softmax warps, which are idle from here to the next tile's first S: the
      // conversion warps go straight on to the next tile's first QKV epilogue -- its projection was issued before this tile's
      // output projection -- instead of sitting ~3000 cycles in the tile-boundary critical path)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc<fattn2::TMEM_COLS>(tmem_base);
  }
}

bool fused_attn2_supported(int D, int heads, int tokens) { return D == fattn2::D && heads == fattn2::NH && tokens == 64; }

// wqkvf: [3D, D] bf16 folded (Q rows scaled by log2(e)/8, all rows times LN gamma); bqp: [D] bf16 folded Q bias;
// wo: [D, D] bf16; bof: [D] fp32 = bo + Wo . (bv + Wv . beta)   (vit_fold.cu)
int launch_fused_attn2(cudaStream_t stream, const __nv_bfloat16* x, __nv_bfloat16* out, const __nv_bfloat16* wqkvf, const __nv_bfloat16* bqp,
                       const __nv_bfloat16* wo, const float* bof, int M, int D, int heads, float eps, const FusedOpts& o) {
  using namespace fattn2;
  if (D != fattn2::D || heads != NH) { set_last_error("fused_attn2: only D=192, 3 heads"); return VITMARL_EINVAL; }
  if (M <= 0) return VITMARL_OK;
  if (M % 64) { set_last_error("fused_attn2: tokens must be whole images of 64"); return VITMARL_EINVAL; }
  CUtensorMap tmX, tmWqkv, tmWo, tmOut;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmX, x, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmOut, out, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmWqkv, wqkvf, 3 * D, D, (uint64_t)D * 2, 64, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmWo, wo, D, D, (uint64_t)D * 2, D, 64))) return rc;
  FusedAttn2Params p{M, x, reinterpret_cast<const uint32_t*>(bqp), bof, eps, o.dbg ? o.dbg + 256 : nullptr, out == x ? 1 : 0, o.attn_flags};
  cudaError_t e = cudaFuncSetAttribute(fused_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e);
  const int tiles = (M + TM - 1) / TM;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(min(tiles, num_sms())); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = o.pdl ? 1 : 0;
  return check_cuda(cudaLaunchKernelEx(&cfg, fused_attn2_kernel, tmX, tmWqkv, tmWo, tmOut, p));
}

}  // namespace vitmarl
