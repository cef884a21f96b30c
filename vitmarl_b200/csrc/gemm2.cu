// 2-CTA variant of the tcgen05 GEMM: a CTA PAIR (thread-block cluster of 2, one TPC) computes a 256 x BN tile with
// `tcgen05.mma.cta_group::2` (M = 256 per instruction).  Each CTA stages its own 128 rows of A and HALF of the B tile
// (BN/2 rows), so every weight byte is fetched once per 256 output rows and the per-instruction issue cost is amortised
// over twice the work of the 1-CTA kernel in gemm.cu.  Operands K-major (activations x weights[out,in]) or MN-major
// (dX = dY . W reads the forward weight buffer as an MN-major B; dW = dY^T . X is MN x MN over the token dimension, split-K).
//
// Protocol (leader = cluster rank 0):
//   * both CTAs: TMA producer loads with `.cta_group::2`, completing transaction bytes on the LEADER's full barrier;
//   * leader only: MMA issuer waits the full barrier, issues the MMAs, `tcgen05.commit...multicast::cluster` arrives on
//     the empty barrier of BOTH CTAs (stage free) and on both accumulator-full barriers;
//   * both CTAs: epilogue warps drain their own TMEM (rows of their CTA) and arrive on the LEADER's accumulator-empty
//     barrier (count = 2 x epilogue warps).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "gemm_epilogue.cuh"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace g2 {
constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BK = 64;
// epilogue warps: 8 for the per-thread epilogue (large K, hidden behind the mainloop), 16 for the staged epilogue of the small-K
// products, whose 8-warp form was longer than the tile's mainloop (gemm_epilogue.cuh)
constexpr int epi_warps(bool tma_epi) { return tma_epi ? 16 : 8; }
constexpr int threads(bool tma_epi) { return 64 + 32 * epi_warps(tma_epi); }
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA
}  // namespace g2

template <int BN, bool TMAEPI, bool B_MN = false>
struct Gemm2Cfg {
  static constexpr int kStages = TMAEPI ? 4 : 6;
  static constexpr int kABytes = g2::BM * g2::BK * 2;          // 16 KB (K-major [128 x 64] or MN-major 2 x [64 k x 64 m])
  // this CTA's half of the B tile.  MN-major: [64 k x 64 n] blocks of 8 KB; BN/2 = 96 columns take one and a half blocks,
  // the second one is loaded whole (its upper 32 columns belong to the peer / the next tile and are never read)
  static constexpr int kBBlocks = (BN / 2 + 63) / 64;
  static constexpr int kBBytes = B_MN ? kBBlocks * 8192 : (BN / 2) * g2::BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = g2::BM * BN * 2;            // bf16 output tile of this CTA, BN/64 swizzled [128 x 64] sub-tiles
  static constexpr int kTmemCols = BN * 2 <= 128 ? 128 : (BN * 2 <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + (TMAEPI ? 2 * kOutBytes : 0) + 1024 + 256;
};

// TMAEPI = true (bf16 outputs): the epilogue goes through shared memory.  A tcgen05.ld hands every thread one ROW of the
// accumulator, so per-thread global loads / stores touch 32 different cache lines per warp instruction (the residual,
// the position embedding and the output all have that shape) and the epilogue -- not the tensor pipe -- bounded the
// small-K projections of the ViT.  Here the addend tile (residual, or a [128, N] bf16 position-embedding tile) is
// TMA-loaded into a swizzled staging tile while the mainloop runs, each thread adds its row segment in place, and the
// tile (plus, for bias+GELU, a second tile with the pre-activation saved for backward) leaves through TMA stores.
template <int BN, bool TMAEPI, bool A_MN = false, bool B_MN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::threads(TMAEPI), 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
             const __grid_constant__ CUtensorMap tmC2, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using namespace g2;
  using Cfg = Gemm2Cfg<BN, TMAEPI, B_MN>;
  constexpr int kEpiWarps = epi_warps(TMAEPI);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = smem_base + Cfg::kStages * Cfg::kStageBytes;                 // TMAEPI: two staging tiles
  const uint32_t bar_base = out_base + (TMAEPI ? 2 * Cfg::kOutBytes : 0);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  auto res_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + 5 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_clusters = gridDim.x >> 1, cluster = blockIdx.x >> 1;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM), tiles_n = p.N / BN;
  const int KB = (p.K + BK - 1) / BK;
  const int num_work = tiles_m * tiles_n * p.splits;     // split-K (EPI_ATOMIC_F32 only): work item = (tile, K range)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * kEpiWarps); mbar_init(res_bar(a), 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int w = cluster; w < num_work; w += num_clusters) {
        const int split = w % p.splits, tile = w / p.splits;
        const int tn = tile % tiles_n, tm = tile / tiles_n;
        const int kb0 = split * p.kb_per_split, kb1 = min(KB, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * Cfg::kStageBytes);      // bytes of BOTH CTAs land on the leader's barrier
          const uint32_t sa = smem_base + s * Cfg::kStageBytes, sb = sa + Cfg::kABytes;
          const uint32_t lbar = full_bar(s) & kPeerMask;
          const int m0 = tm * 2 * BM + (int)rank * BM, n0 = tn * BN + (int)rank * (BN / 2);
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, kb * BK, m0, lbar);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d_2sm(sa + i * 8192, &tmA, m0 + i * 64, kb * BK, lbar);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmB, kb * BK, n0, lbar);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::kBBlocks; ++i) tma_load_2d_2sm(sb + i * 8192, &tmB, n0 + i * 64, kb * BK, lbar);
          }
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only; whole warp loops, one elected lane issues) =================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int w = cluster; w < num_work; w += num_clusters, ++it) {
        const int split = w % p.splits;
        const int kb0 = split * p.kb_per_split, kb1 = min(KB, kb0 + p.kb_per_split);
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * Cfg::kStageBytes, sb = sa + Cfg::kABytes;
          // MN-major operand: 64-element blocks 8192 B apart (LBO), a K step of 16 rows = +2048 B; K-major: +32 B inside the swizzled row
          const uint32_t la = umma_desc_lo(sa, A_MN ? 8192 : 16), lb = umma_desc_lo(sb, B_MN ? 8192 : 16);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2sm(d_tmem, umma_desc_from_lo(la + k * (A_MN ? 128 : 2)), umma_desc_from_lo(lb + k * (B_MN ? 128 : 2)), idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(empty_bar(s));
          }
          __syncwarp();
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2sm(tfull_bar(acc));
        __syncwarp();
      }
    }
  } else if constexpr (!TMAEPI) {
    // ================= epilogue (both CTAs; warps 2..9): per-thread global accesses =================
    const int quad = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
    int it = 0;
    for (int w = cluster; w < num_work; w += num_clusters, ++it) {
      const int tile = w / p.splits;
      const int tn = tile % tiles_n, tm = tile / tiles_n;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const int row = tm * 2 * BM + (int)rank * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = chalf * kColsPerWarp; c < (chalf + 1) * kColsPerWarp; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        gemm_epilogue_chunk(p, r, row, row_ok, tn * BN + c);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc) & kPeerMask);
    }
  } else {
    // ================= epilogue (both CTAs; warps 2..17): staged in shared memory, TMA in / out (gemm_epilogue.cuh) =================
    extern __shared__ uint8_t smem_gen[];
    uint8_t* sgen = smem_gen + (smem_base - smem_u32(smem_gen));          // generic pointer to the aligned base
    int it = 0;
    for (int w = cluster; w < num_work; w += num_clusters, ++it) {      // (splits == 1 on this path)
      const int tn = w % tiles_n, tm = w / tiles_n;
      const int acc = it & 1;
      const uint32_t tbar = tempty_bar(acc) & kPeerMask;
      staged_epilogue_tile<BN, kEpiWarps>(p, &tmC, &tmC2, &tmR, sgen, smem_base, out_base, res_bar(0), tfull_bar(acc), (it >> 1) & 1,
                               tmem_base + acc * BN, it, tn * BN, tm * 2 * BM + (int)rank * BM, warp, lane,
                               [&] { mbar_arrive_cluster(tbar); }, w + num_clusters < num_work, ((w + num_clusters) % tiles_n) * BN,
                               ((w + num_clusters) / tiles_n) * 2 * BM + (int)rank * BM);
    }
    if (warp == 2 && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // neither CTA exits while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
  }
}

bool gemm2_supported(const GemmDesc& g) {
  if (!g.allow_2cta || g.M < 512 || g.N % 192) return false;
  if (!g.a_mn_major && !g.b_mn_major) return g.epi != EPI_ATOMIC_F32;
  if (!g.a_mn_major && g.b_mn_major) return g.epi != EPI_ATOMIC_F32;             // dX = dY . W (W read as an MN-major B)
  // dW = dY^T . X: both MN-major, split-K with fp32 red.add; output rows = features, worth it when 256-row pair tiles fill up
  if (g.a_mn_major && g.b_mn_major) return g.epi == EPI_ATOMIC_F32 && (g.M % 256 == 0 || g.M >= 1024);
  return false;
}

template <bool TMAEPI, bool A_MN, bool B_MN>
static int launch_gemm2_t(cudaStream_t stream, const GemmDesc& g, int add_mode, const __nv_bfloat16* add, int add_rows, int add_ld) {
  using namespace g2;
  constexpr int BN = 192;
  using Cfg = Gemm2Cfg<BN, TMAEPI, B_MN>;
  CUtensorMap tmA, tmB, tmC, tmC2, tmR;
  int rc;
  // K-major operand: tensor [rows = M or N, cols = K]; MN-major operand: tensor [rows = K, cols = M or N], box [64 rows, 64 cols]
  if (!A_MN) rc = make_tmap_2d_bf16(&tmA, g.A, g.M, g.K, (uint64_t)g.lda * 2, BM, BK);
  else rc = make_tmap_2d_bf16(&tmA, g.A, g.K, g.M, (uint64_t)g.lda * 2, BK, 64);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d_bf16(&tmB, g.B, g.N, g.K, (uint64_t)g.ldb * 2, BN / 2, BK);
  else rc = make_tmap_2d_bf16(&tmB, g.B, g.K, g.N, (uint64_t)g.ldb * 2, BK, 64);
  if (rc) return rc;
  tmC = tmA; tmC2 = tmA; tmR = tmA;                      // placeholders when unused
  if (TMAEPI) {
    if ((rc = make_tmap_2d_bf16(&tmC, g.C, g.M, g.N, (uint64_t)g.ldc * 2, BM, 64))) return rc;
    if (g.C2 && (rc = make_tmap_2d_bf16(&tmC2, g.C2, g.M, g.N, (uint64_t)g.ldc * 2, BM, 64))) return rc;
    if (add_mode && (rc = make_tmap_2d_bf16(&tmR, add, add_rows, g.N, (uint64_t)add_ld * 2, BM, 64))) return rc;
  }
  GemmParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  const int KB = (g.K + BK - 1) / BK;
  const int tiles = ((g.M + 2 * BM - 1) / (2 * BM)) * (g.N / BN);
  const int max_clusters = num_sms() / 2;
  int splits = 1;
  if (g.epi == EPI_ATOMIC_F32) splits = max(1, min(KB, (2 * max_clusters + tiles - 1) / tiles));
  p.kb_per_split = (KB + splits - 1) / splits;
  p.splits = (KB + p.kb_per_split - 1) / p.kb_per_split;
  p.epi = g.epi; p.C = g.C; p.C2 = g.C2; p.ldc = g.ldc; p.bias = g.bias; p.residual = g.residual; p.ldr = g.ldr;
  p.pos = g.pos; p.pos_period = g.pos_period > 0 ? g.pos_period : 1; p.out_scale = g.out_scale;
  p.add_mode = TMAEPI ? add_mode : 0;
  auto kern = gemm2_kernel<BN, TMAEPI, A_MN, B_MN>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (err != cudaSuccess) return check_cuda(err);
  const int work = tiles * p.splits;
  const int clusters = min(work, max_clusters);
  kern<<<2 * clusters, threads(TMAEPI), Cfg::kSmemBytes, stream>>>(tmA, tmB, tmC, tmC2, tmR, p);
  return check_cuda(cudaGetLastError());
}

int launch_gemm2(cudaStream_t stream, const GemmDesc& g) {
  if (g.a_mn_major) return launch_gemm2_t<false, true, true>(stream, g, 0, nullptr, 0, 0);      // dW: split-K, fp32 red.add epilogue
  const __nv_bfloat16* add; int add_rows, add_ld;
  const int mode = staged_epilogue_mode(g, &add, &add_rows, &add_ld);
  // small-K products are epilogue-bound (staged epilogue, 4-stage ring); large-K products are mainloop-bound and keep the
  // 6-stage ring with the per-thread epilogue hidden behind the next tile's mainloop
  const bool staged = mode >= 0 && g.K <= 1536;   // (also measured for the dX shapes at D = 384: the deeper ring does not beat it)
  if (g.b_mn_major) {
    if (staged) return launch_gemm2_t<true, false, true>(stream, g, mode, add, add_rows, add_ld);
    return launch_gemm2_t<false, false, true>(stream, g, 0, nullptr, 0, 0);
  }
  if (staged) return launch_gemm2_t<true, false, false>(stream, g, mode, add, add_rows, add_ld);
  return launch_gemm2_t<false, false, false>(stream, g, 0, nullptr, 0, 0);
}

}  // namespace vitmarl
