// Weight-gradient GEMM on CTA pairs:  dW[M, N] += scale * A^T . B  over the token dimension K, both operands MN-major
// (A = dY [K tokens, M features], B = X [K tokens, N features]), split-K with fp32 red.add.
//
// Why a kernel of its own: this product streams BOTH operands (nothing is a resident weight), and with ~2100 cycles of TMA
// latency the shared-memory ring (bytes in flight / latency) caps the feed at ~90 B/clk per SM.  A 256 x 192 pair tile needs
// 85 B/clk at full tensor rate and ran at half of it; a 256 x 384 work item (one A tile against 384 columns, two MMAs per
// K step: N = 256 and N = 128, both 64-column-block aligned) needs 53 B/clk.  The accumulator (384 of 512 TMEM columns)
// is single-buffered: with split-K a work item is a >100 us main loop and a ~1 us epilogue.
//
// Per CTA and stage: A 2 x [64 k x 64 m] blocks (its 128 rows), B 3 x [64 k x 64 n] blocks (its 192 of the 384 columns).
// Accumulator columns -> output columns: MMA #0 (N = 256) takes blocks 0-1 of each CTA, MMA #1 (N = 128) block 2:
//   acc [  0,128) -> n0 +   0 .. 127      acc [128,256) -> n0 + 192 .. 319      (CTA 0 / CTA 1 blocks 0-1)
//   acc [256,320) -> n0 + 128 .. 191      acc [320,384) -> n0 + 320 .. 383      (CTA 0 / CTA 1 block 2)
// `transpose_out` writes C[col][row] instead (lets the caller swap the operands when M is the short dimension).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace g2dw {
constexpr int BM = 128, BK = 64, NT = 384;
constexpr int kStages = 5;
constexpr int kABytes = 2 * 8192, kBBytes = 3 * 8192, kStageBytes = kABytes + kBBytes;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kTmemCols = 512;
constexpr int kOnesBytes = 2048;   // [16 k x 64 n] bf16 ones: B operand of the column-sum MMA
constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 + 256;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
}  // namespace g2dw

struct DwParams {
  int M, N, K, kb_per_split, splits;
  float* C; int ldc; float scale; int transpose_out;
  float* colsum;   // optional: colsum[m] += sum_k A[k][m] (the bias gradient of the layer whose dW this is)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2dw::kThreads, 1)
gemm2_dw_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DwParams p) {
  using namespace g2dw;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ones_base = smem_base + kStages * kStageBytes;
  const uint32_t bar_base = ones_base + kOnesBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kStages), tempty_bar = bar_base + 8u * (2 * kStages + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_clusters = gridDim.x >> 1, cluster = blockIdx.x >> 1;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM), tiles_n = p.N / NT;
  const int KB = (p.K + BK - 1) / BK;
  const int num_work = tiles_m * tiles_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 2 * kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (p.colsum) {                                         // ones tile (generic-proxy writes -> visible to the tensor core's async proxy)
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += kThreads)
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(ones_base + 4u * i), "r"(0x3F803F80u) : "memory");
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
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
        const int m0 = tm * 2 * BM + (int)rank * BM, n0 = tn * NT + (int)rank * (NT / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * kStageBytes);
          const uint32_t sa = smem_base + s * kStageBytes, sb = sa + kABytes;
          const uint32_t lbar = full_bar(s) & kPeerMask;
#pragma unroll
          for (int i = 0; i < 2; ++i) tma_load_2d_2sm(sa + i * 8192, &tmA, m0 + i * 64, kb * BK, lbar);
#pragma unroll
          for (int i = 0; i < 3; ++i) tma_load_2d_2sm(sb + i * 8192, &tmB, n0 + i * 64, kb * BK, lbar);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      constexpr uint32_t id256 = umma_idesc_bf16(2 * BM, 256, true, true), id128 = umma_idesc_bf16(2 * BM, 128, true, true);
      constexpr uint32_t id16 = umma_idesc_bf16(2 * BM, 16, true, true);     // A^T . ones: 16 identical columns of column sums
      const uint64_t d_ones = umma_desc_from_lo(umma_desc_lo(ones_base, 8192));
      const bool do_colsum = p.colsum != nullptr;
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int w = cluster; w < num_work; w += num_clusters, ++it) {
        const int split = w % p.splits;
        const bool cs = do_colsum && ((w / p.splits) % tiles_n) == 0;    // the first column tile of each row block carries the sums
        const int kb0 = split * p.kb_per_split, kb1 = min(KB, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar, (it & 1) ^ 1);             // previous work item drained from tensor memory
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * kStageBytes, sb = sa + kABytes;
          const uint32_t la = umma_desc_lo(sa, 8192), lb = umma_desc_lo(sb, 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              const uint64_t da = umma_desc_from_lo(la + k * 128);
              umma_bf16_2sm(tmem_base, da, umma_desc_from_lo(lb + k * 128), id256, accum);                         // blocks 0-1
              umma_bf16_2sm(tmem_base + 256, da, umma_desc_from_lo(lb + (2 * 8192 >> 4) + k * 128), id128, accum);  // block 2
              if (cs) umma_bf16_2sm(tmem_base + NT, da, d_ones, id16, accum);
            }
            umma_commit_2sm(empty_bar(s));
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2sm(tfull_bar);
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue (both CTAs; warps 2..9): fp32 red.add =================
    const int quad = warp & 3;
    const int chalf = (warp - 2) >> 2;
    int it = 0;
    for (int w = cluster; w < num_work; w += num_clusters, ++it) {
      const int tile = w / p.splits;
      const int tn = tile % tiles_n, tm = tile / tiles_n;
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
      const int row = tm * 2 * BM + (int)rank * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c = chalf * (NT / 2); c < (chalf + 1) * (NT / 2); c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
        const int col = tn * NT + (c < 256 ? (c >> 7) * 192 + (c & 127) : ((c - 256) >> 6) * 192 + 128 + ((c - 256) & 63));
        if (row_ok) {
          if (!p.transpose_out) {
            float* dst = p.C + (size_t)row * p.ldc + col;   // 32 consecutive floats of one output row: 16-byte vector reductions
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j]) * p.scale),
                           "f"(__uint_as_float(r[j + 1]) * p.scale), "f"(__uint_as_float(r[j + 2]) * p.scale),
                           "f"(__uint_as_float(r[j + 3]) * p.scale) : "memory");
          } else {
            float* dst = p.C + (size_t)col * p.ldc + row;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * p.ldc, __uint_as_float(r[j]) * p.scale);
          }
        }
      }
      if (p.colsum && tn == 0 && chalf == 0) {             // column NT of the accumulator: sum over this K range of A[k][row]
        uint32_t r16[16];
        tmem_ld_32x16(taddr + NT, r16);
        tmem_ld_wait();
        if (row_ok) atomicAdd(p.colsum + row, __uint_as_float(r16[0]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar & kPeerMask);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// A = g.A viewed [K, M] (MN-major), B = g.B viewed [K, N]; C fp32 [M, N] (or [N, M] when transpose_out), ld = g.ldc
static int launch_dw_pair(cudaStream_t stream, const __nv_bfloat16* A, int lda, int M, const __nv_bfloat16* B, int ldb, int N, int K,
                          float* C, int ldc, float scale, bool transpose_out, float* colsum) {
  using namespace g2dw;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmA, A, K, M, (uint64_t)lda * 2, BK, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, B, K, N, (uint64_t)ldb * 2, BK, 64))) return rc;
  DwParams p{};
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.scale = scale; p.transpose_out = transpose_out ? 1 : 0; p.colsum = colsum;
  const int KB = (K + BK - 1) / BK;
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / NT);
  const int max_clusters = num_sms() / 2;
  // one round of work items (the accumulator is single-buffered and every item ends in 98 K reductions): as many
  // K ranges as fill the clusters once
  const int splits = max(1, min(KB, max_clusters / tiles));
  p.kb_per_split = (KB + splits - 1) / splits;
  p.splits = (KB + p.kb_per_split - 1) / p.kb_per_split;
  cudaError_t err = cudaFuncSetAttribute(gemm2_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (err != cudaSuccess) return check_cuda(err);
  const int clusters = min(tiles * p.splits, max_clusters);
  gemm2_dw_kernel<<<2 * clusters, kThreads, kSmemBytes, stream>>>(tmA, tmB, p);
  return check_cuda(cudaGetLastError());
}

// Takes dW-shaped products (both operands MN-major, fp32 accumulate-into) whose output has a 384-multiple dimension and a long
// other dimension; returns 1 when the shape is not its business (the caller falls back to the generic kernels).
int launch_gemm2_dw(cudaStream_t stream, const GemmDesc& g) {
  if (!(g.a_mn_major && g.b_mn_major && g.epi == EPI_ATOMIC_F32) || g.M % 64 || g.N % 64 || g.ldc % 4 ||
      (reinterpret_cast<uintptr_t>(g.C) & 15)) return 1;
  // rows of the 256-row pair tiles that do real work, in 1/1024ths; the short side may waste up to 25 %
  auto fill = [](int m) { return (int)(1024LL * m / ((m + 255) / 256 * 256)); };
  float* C = reinterpret_cast<float*>(g.C);
  const int direct = g.N % g2dw::NT == 0 ? fill(g.M) : 0;      // C   = A^T . B
  const int swapped = g.M % g2dw::NT == 0 ? fill(g.N) : 0;     // C^T = B^T . A: the other dimension becomes the row dimension of the tiles
  if (max(direct, swapped) < 768) return 1;
  if (swapped > direct || (swapped == direct && !g.colsum_a)) {    // (a fused bias gradient needs A on the row side)
    if (g.colsum_a) { const int rc = launch_colsum(stream, g.A, g.colsum_a, g.K, g.M); if (rc) return rc; }
    return launch_dw_pair(stream, g.B, g.ldb, g.N, g.A, g.lda, g.M, g.K, C, g.ldc, g.out_scale, true, nullptr);
  }
  return launch_dw_pair(stream, g.A, g.lda, g.M, g.B, g.ldb, g.N, g.K, C, g.ldc, g.out_scale, false, g.colsum_a);
}

}  // namespace vitmarl
