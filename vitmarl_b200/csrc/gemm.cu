// bf16 GEMM on the 5th-gen tensor cores: TMA (128B swizzle) -> smem ring -> tcgen05.mma
// (accumulators in TMEM, double buffered) -> tcgen05.ld epilogue with fused bias / GELU /
// residual / positional-embedding add.  Persistent, warp specialised:
//   warp 0   TMA producer (one elected lane)
//   warp 1   TMEM allocator + MMA issuer (one elected lane)
//   warps 2-9 epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 (one output row per thread) and half of
//             the tile's columns (the epilogue -- TMEM loads, bias/GELU math, stores -- is the per-SM
//             issue bottleneck for these skinny-K products, so it gets 8 of the 10 warps)
//
//   C[M,N] = epi( A[M,K] . B[N,K]^T )        A, B bf16; fp32 accumulate
// A / B may each be "K-major" (contraction index contiguous: activations [tokens, features],
// weights [out, in]) or "MN-major" (the operand is stored transposed, e.g. dW = dY^T X where
// both operands have the token index as the slow dimension).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "gemm_epilogue.cuh"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8;                       // 2 warps per TMEM lane quadrant, each owns half of the tile's columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;

template <int BN, bool STAGED>
struct GemmCfg {
  // STAGED: two bf16 staging tiles for the shared-memory epilogue (gemm_epilogue.cuh) take the place of ring stages
  static constexpr int kStages = STAGED ? (BN >= 192 ? 3 : (BN >= 128 ? 4 : 6)) : (BN >= 192 ? 5 : (BN >= 128 ? 6 : 8));
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBytes = BM * BN * 2;
  static constexpr int kTmemCols = BN * 2 <= 128 ? 128 : (BN * 2 <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + (STAGED ? 2 * kOutBytes : 0) + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN, bool STAGED>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
            const __grid_constant__ CUtensorMap tmC2, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using Cfg = GemmCfg<BN, STAGED>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = smem_base + Cfg::kStages * Cfg::kStageBytes;                 // STAGED: two staging tiles
  const uint32_t bar_base = out_base + (STAGED ? 2 * Cfg::kOutBytes : 0);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  auto res_bar = [&](int b) { return bar_base + 8u * (2 * Cfg::kStages + 5 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = p.N / BN;
  const int KB = (p.K + BK - 1) / BK;
  const int num_work = tiles_m * tiles_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiWarps); mbar_init(res_bar(a), 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int split = w % p.splits, tile = w / p.splits;
        const int tn = tile % tiles_n, tm = tile / tiles_n;
        const int kb0 = split * p.kb_per_split, kb1 = min(KB, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), Cfg::kStageBytes);
          const uint32_t sa = smem_base + s * Cfg::kStageBytes, sb = sa + Cfg::kABytes;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, kb * BK, tm * BM, full_bar(s));
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * 8192, &tmA, tm * BM + i * 64, kb * BK, full_bar(s));
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, kb * BK, tn * BN, full_bar(s));
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * 8192, &tmB, tn * BN + i * 64, kb * BK, full_bar(s));
          }
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the loop and one elected lane issues (a plain predicated instruction stream; under
    // `if (lane == 0)` the compiler wraps every tcgen05.mma in a per-active-thread loop); descriptors are built once per
    // stage and advanced by immediates (+32 B along K inside a swizzled row = +2, +2048 B for an MN-major K step = +128).
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
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
          const uint32_t la = umma_desc_lo(sa, A_MN ? 8192 : 16), lb = umma_desc_lo(sb, B_MN ? 8192 : 16);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(d_tmem, umma_desc_from_lo(la + k * (A_MN ? 128 : 2)), umma_desc_from_lo(lb + k * (B_MN ? 128 : 2)), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(empty_bar(s));                   // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tfull_bar(acc));    // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else if constexpr (STAGED) {
    // ================= epilogue (warps 2..9): staged in shared memory, TMA in / out (gemm_epilogue.cuh; never split-K) =================
    extern __shared__ uint8_t smem_gen[];
    uint8_t* sgen = smem_gen + (smem_base - smem_u32(smem_gen));
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int tn = w % tiles_n, tm = w / tiles_n;
      const int acc = it & 1;
      const uint32_t tbar = tempty_bar(acc);
      staged_epilogue_tile<BN>(p, &tmC, &tmC2, &tmR, sgen, smem_base, out_base, res_bar(0), tfull_bar(acc), (it >> 1) & 1,
                               tmem_base + acc * BN, it, tn * BN, tm * BM, warp, lane, [&] { mbar_arrive(tbar); },
                               w + (int)gridDim.x < num_work, ((w + (int)gridDim.x) % tiles_n) * BN, ((w + (int)gridDim.x) / tiles_n) * BM);
    }
    if (warp == 2 && lane == 0) bulk_wait0();
  } else {
    // ================= epilogue (warps 2..9) =================
    const int quad = warp & 3;
    const int chalf = (warp - 2) >> 2;           // which half of the columns
    constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int tile = w / p.splits;
      const int tn = tile % tiles_n, tm = tile / tiles_n;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const int row = tm * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = chalf * kColsPerWarp; c < (chalf + 1) * kColsPerWarp; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        const int col = tn * BN + c;
        gemm_epilogue_chunk(p, r, row, row_ok, col);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------- host
static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, bool A_MN, bool B_MN, bool STAGED>
static int launch_gemm_t2(cudaStream_t stream, const GemmDesc& g, int add_mode, const __nv_bfloat16* add, int add_rows, int add_ld) {
  using Cfg = GemmCfg<BN, STAGED>;
  CUtensorMap tmA, tmB, tmC, tmC2, tmR;
  int rc;
  // K-major operand: tensor [rows = M or N, cols = K], box [BM or BN rows, 64 cols]
  // MN-major operand: tensor [rows = K, cols = M or N], box [64 rows, 64 cols]
  if (!A_MN) rc = make_tmap_2d_bf16(&tmA, g.A, g.M, g.K, (uint64_t)g.lda * 2, BM, BK);
  else rc = make_tmap_2d_bf16(&tmA, g.A, g.K, g.M, (uint64_t)g.lda * 2, BK, 64);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d_bf16(&tmB, g.B, g.N, g.K, (uint64_t)g.ldb * 2, BN, BK);
  else rc = make_tmap_2d_bf16(&tmB, g.B, g.K, g.N, (uint64_t)g.ldb * 2, BK, 64);
  if (rc) return rc;
  tmC = tmA; tmC2 = tmA; tmR = tmA;                      // placeholders when unused
  if (STAGED) {
    if ((rc = make_tmap_2d_bf16(&tmC, g.C, g.M, g.N, (uint64_t)g.ldc * 2, BM, 64))) return rc;
    if (g.C2 && (rc = make_tmap_2d_bf16(&tmC2, g.C2, g.M, g.N, (uint64_t)g.ldc * 2, BM, 64))) return rc;
    if (add_mode && (rc = make_tmap_2d_bf16(&tmR, add, add_rows, g.N, (uint64_t)add_ld * 2, BM, 64))) return rc;
  }
  GemmParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  const int KB = (g.K + BK - 1) / BK;
  const int tiles = ((g.M + BM - 1) / BM) * (g.N / BN);
  int splits = 1;
  if (g.epi == EPI_ATOMIC_F32) {
    splits = max(1, min(KB, (2 * num_sms() + tiles - 1) / tiles));
  }
  p.kb_per_split = (KB + splits - 1) / splits;
  p.splits = (KB + p.kb_per_split - 1) / p.kb_per_split;
  p.epi = g.epi; p.C = g.C; p.C2 = g.C2; p.ldc = g.ldc; p.bias = g.bias; p.residual = g.residual; p.ldr = g.ldr;
  p.pos = g.pos; p.pos_period = g.pos_period > 0 ? g.pos_period : 1; p.out_scale = g.out_scale;
  p.add_mode = STAGED ? add_mode : 0;
  auto kern = gemm_kernel<BN, A_MN, B_MN, STAGED>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (err != cudaSuccess) return check_cuda(err);
  const int work = tiles * p.splits;
  const int grid = min(work, num_sms());
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmC, tmC2, tmR, p);
  return check_cuda(cudaGetLastError());
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm_t(cudaStream_t stream, const GemmDesc& g) {
  const __nv_bfloat16* add; int add_rows, add_ld;
  const int mode = staged_epilogue_mode(g, &add, &add_rows, &add_ld);
  // the staged epilogue pays off for the epilogue-bound token-dimension GEMMs with a small contraction; others keep the deeper ring
  if (mode >= 0 && g.M >= 1024 && g.K <= 1536) return launch_gemm_t2<BN, A_MN, B_MN, true>(stream, g, mode, add, add_rows, add_ld);
  return launch_gemm_t2<BN, A_MN, B_MN, false>(stream, g, 0, nullptr, 0, 0);
}

template <int BN>
static int launch_gemm_bn(cudaStream_t stream, const GemmDesc& g) {
  if (!g.a_mn_major && !g.b_mn_major) return launch_gemm_t<BN, false, false>(stream, g);
  if (g.a_mn_major && g.b_mn_major) return launch_gemm_t<BN, true, true>(stream, g);
  if (!g.a_mn_major && g.b_mn_major) return launch_gemm_t<BN, false, true>(stream, g);
  return launch_gemm_t<BN, true, false>(stream, g);
}

int launch_gemm(cudaStream_t stream, const GemmDesc& g) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0 || !g.A || !g.B || !g.C) { set_last_error("gemm: bad arguments"); return VITMARL_EINVAL; }
  if (g.N % 64 || g.K % 8 || g.lda % 8 || g.ldb % 8 || (g.a_mn_major && g.M % 64)) { set_last_error("gemm: unsupported shape"); return VITMARL_EINVAL; }
  if (g.allow_2cta && g.a_mn_major && g.b_mn_major && g.epi == EPI_ATOMIC_F32) {
    const int rc = launch_gemm2_dw(stream, g);           // (fuses g.colsum_a when it runs the product un-swapped)
    if (rc != 1) return rc;                              // 1: not a shape for the 256 x 384 weight-gradient kernel
  }
  if (g.colsum_a) {                                      // generic kernels: the bias gradient is a pass of its own
    if (!g.a_mn_major) { set_last_error("gemm: colsum_a needs an MN-major A"); return VITMARL_EINVAL; }
    const int rc = launch_colsum(stream, g.A, g.colsum_a, g.K, g.M);
    if (rc) return rc;
  }
  if (gemm2_supported(g)) return launch_gemm2(stream, g);
  if (g.N % 192 == 0) return launch_gemm_bn<192>(stream, g);
  if (g.N % 128 == 0) return launch_gemm_bn<128>(stream, g);
  return launch_gemm_bn<64>(stream, g);
}

}  // namespace vitmarl
