// Shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels: kernel parameters and the fused epilogue applied to one
// 32-column chunk of one accumulator row (the registers come from tcgen05.ld 32x32b.x32).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"

namespace vitmarl {

struct GemmParams {
  int M, N, K;            // GEMM extents (K = contraction)
  int kb_per_split;       // k-blocks per split-K slice
  int splits;
  int epi;                // GemmEpi
  void* C; int ldc;       // bf16 or fp32 (atomic) output, row pitch in elements
  void* C2;               // EPI_BIAS_GELU: optional bf16 copy of the pre-activation (saved for backward)
  const float* bias;      // [N] or null
  const __nv_bfloat16* residual; int ldr;   // [M, N] bf16 or null
  const float* pos; int pos_period;         // [pos_period, N] fp32 (row % pos_period) or null
  float out_scale;
  int add_mode;           // shared-memory-staged epilogue: 0 no addend tile, 1 residual tile (+=), 2 position-embedding tile (+=, row 0),
                          // 3 pre-activation tile (*= gelu_tanh'(aux): EPI_MUL_GELU_GRAD)
};


__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, uint32_t (&r)[32], int row, bool row_ok, int col) {
  if (p.epi == EPI_ATOMIC_F32) {
    if (row_ok) {
      float* dst = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {      // 32 consecutive floats of one row: 16-byte vector reductions
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j]) * p.out_scale),
                       "f"(__uint_as_float(r[j + 1]) * p.out_scale), "f"(__uint_as_float(r[j + 2]) * p.out_scale),
                       "f"(__uint_as_float(r[j + 3]) * p.out_scale) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(r[j]) * p.out_scale);
      }
    }
  } else {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (p.epi == EPI_BIAS_GELU) {
      if (p.C2 && row_ok) {
        uint4* dst2 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C2) + (size_t)row * p.ldc + col);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst2[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                               pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
    }
    if (row_ok) {
      if (p.pos) {
        const float* pp = p.pos + (size_t)(row % p.pos_period) * p.N + col;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 b = __ldg(reinterpret_cast<const float4*>(pp + j));
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
      if (p.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.ldr + col);
        float a[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q = __ldg(rp + j);
          a[8 * j + 0] = bf16_lo(q.x); a[8 * j + 1] = bf16_hi(q.x);
          a[8 * j + 2] = bf16_lo(q.y); a[8 * j + 3] = bf16_hi(q.y);
          a[8 * j + 4] = bf16_lo(q.z); a[8 * j + 5] = bf16_hi(q.z);
          a[8 * j + 6] = bf16_lo(q.w); a[8 * j + 7] = bf16_hi(q.w);
        }
        if (p.epi == EPI_MUL_GELU_GRAD) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= gelu_tanh_grad(a[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += a[j];
        }
      }
      if (p.epi == EPI_STORE_F32) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                              pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
      }
    }
  }
}

// ---- shared-memory-staged epilogue of one [128 x BN] bf16 output tile (NW epilogue warps = 32 NW threads, named barrier 1) ----
// A tcgen05.ld hands every thread one ROW of the accumulator, so per-thread global loads / stores touch 32 different cache
// lines per warp instruction (the residual, the pre-activation, the position embedding and the output all have that
// shape) and the epilogue -- not the tensor pipe -- bounded the small-K projections of the ViT.  Here the addend tile is
// TMA-loaded into a swizzled staging tile while the mainloop of the same tile runs, each thread combines its row segment
// in place, and the tile (plus, for bias+GELU, a second tile with the pre-activation saved for backward) leaves through
// TMA stores.  Staging tiles: BN/64 sub-tiles of [128 rows x 64 cols], 128B swizzle (chunk ^ (row & 7)).
//   out_base : two staging tiles of 128*BN*2 bytes (alternating per tile; C and C2 when both are written)
//   res_bar0 : two mbarriers (one per staging tile) for the addend loads
// NW = 8: two warps per TMEM lane quadrant, BN/2 columns each in 32-column chunks.  NW = 16 (the CTA-pair kernel): FOUR warps per
// quadrant, BN/4 columns each in 16-column chunks -- at K = 384 the 8-warp epilogue (~5.5 k cycles per tile, issue-active 25 %:
// two warps per scheduler stalled on TMEM / shared-memory / bias loads) was more than twice the tile's 2.3 k-cycle mainloop.
template <int BN, int NW = 8, typename ArriveEmpty>
__device__ __forceinline__ void staged_epilogue_tile(const GemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmC2,
                                                     const CUtensorMap* tmR, uint8_t* sgen, uint32_t smem_base, uint32_t out_base,
                                                     uint32_t res_bar0, uint32_t tfull_bar, uint32_t acc_ph, uint32_t taddr_acc,
                                                     int it, int col0, int grow0, int warp, int lane, ArriveEmpty arrive_empty,
                                                     bool has_next = false, int next_col0 = 0, int next_grow0 = 0) {
  constexpr int kOutBytes = 128 * BN * 2;
  constexpr int kColsPerWarp = BN / (NW / 4);
  constexpr int CW = (kColsPerWarp % 32 == 0) ? 32 : 16;             // columns per tcgen05.ld
  static_assert(kColsPerWarp % CW == 0 && (NW == 8 || NW == 16), "epilogue warp layout");
  const int quad = warp & 3, cpart = (warp - 2) >> 2;
  const int trow = quad * 32 + lane;
  const uint32_t sw = (uint32_t)(trow & 7);
  const bool has_add = p.add_mode != 0;
  const bool two_out = p.epi == EPI_BIAS_GELU && p.C2 != nullptr;
  const bool elected = (warp == 2 && lane == 0);
  const int buf = two_out ? 0 : (it & 1);
  const uint32_t res_ph = two_out ? (it & 1) : ((it >> 1) & 1);
  const uint32_t out0 = out_base + buf * kOutBytes, out1 = out_base + kOutBytes;
  const uint32_t res_bar = res_bar0 + 8u * buf;
  auto load_addend = [&](uint32_t dst, uint32_t bar_, int c0, int g0) {
    mbar_arrive_expect_tx(bar_, kOutBytes);
    for (int kb = 0; kb < BN / 64; ++kb) tma_load_2d(dst + kb * (128 * 128), tmR, c0 + kb * 64, p.add_mode == 2 ? 0 : g0, bar_);
  };
  auto epi_barrier = [&]() {
    if (NW == 8) asm volatile("bar.sync 1, 256;" ::: "memory");
    else asm volatile("bar.sync 1, 512;" ::: "memory");
  };
  if (elected) {
    // the staging tile(s) of this iteration must have been read by their previous TMA store
    if (two_out) bulk_wait_read0();
    else if (!has_add) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    // addend tile: requested ONE TILE AHEAD (at the end of the previous tile's epilogue, below) so that its ~2000 cycles of
    // TMA latency run under the previous epilogue instead of in front of this one; only the first tile requests its own
    if (has_add && it == 0) load_addend(out0, res_bar, col0, grow0);
  }
  epi_barrier();
  mbar_wait(tfull_bar, acc_ph);
  if (has_add) mbar_wait(res_bar, res_ph);
  tc_fence_after();
  const uint32_t taddr = taddr_acc + ((uint32_t)(quad * 32) << 16);
  uint8_t* o0 = sgen + (out0 - smem_base);
  uint8_t* o1 = sgen + (out1 - smem_base);
#pragma unroll 1
  for (int c = cpart * kColsPerWarp; c < (cpart + 1) * kColsPerWarp; c += CW) {
    uint32_t r[CW];
    if constexpr (CW == 32) tmem_ld_32x32(taddr + c, r);
    else tmem_ld_32x16(taddr + c, r);
    tmem_ld_wait();
    if (c + CW >= (cpart + 1) * kColsPerWarp) {                            // last read of the accumulator by this warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_empty();
    }
    float v[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c + j));   // same address in every lane
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    const uint32_t tile_off = (uint32_t)((c >> 6) * (128 * 128) + trow * 128);
    const uint32_t ch0 = (uint32_t)((c & 63) >> 3);
    if (p.epi == EPI_BIAS_GELU) {
      // GELU on PACKED bf16 pairs of the (bf16-rounded) pre-activation -- the value the backward pass sees and the arithmetic the
      // fused inference block uses: 6 packed instructions per 2 elements instead of ~12 fp32 instructions per element.  With K = 384
      // the fp32 form made this epilogue (3 k issue cycles per 128 x 192 tile on 8 warps) longer than the tile's mainloop (2.3 k).
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) {
        const uint4 pre = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                     pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        if (two_out) *reinterpret_cast<uint4*>(o1 + tile_off + (((ch0 + j) ^ sw) << 4)) = pre;
        *reinterpret_cast<uint4*>(o0 + tile_off + (((ch0 + j) ^ sw) << 4)) =
            make_uint4(gelu_tanh_bf16x2(pre.x), gelu_tanh_bf16x2(pre.y), gelu_tanh_bf16x2(pre.z), gelu_tanh_bf16x2(pre.w));
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) {
      uint4* slot = reinterpret_cast<uint4*>(o0 + tile_off + (((ch0 + j) ^ sw) << 4));
      if (has_add) {
        const uint4 a = *slot;
        const float a8[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y), bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
        if (p.add_mode == 3) {
          // (fp32 on purpose: gelu' on packed bf16 pairs was measured -- 1 % faster on ViT-S/16 -- but 1 + tanh cancels for negative
          // pre-activations and the element-wise error reached 2.8 % of the output range)
#pragma unroll
          for (int k = 0; k < 8; ++k) v[8 * j + k] *= gelu_tanh_grad(a8[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[8 * j + k] += a8[k];
        }
      }
      *slot = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    }
  }
  fence_proxy_async_smem();
  epi_barrier();
  if (elected) {
    for (int kb = 0; kb < BN / 64; ++kb) tma_store_2d(tmC, out0 + kb * (128 * 128), col0 + kb * 64, grow0);
    if (two_out)
      for (int kb = 0; kb < BN / 64; ++kb) tma_store_2d(tmC2, out1 + kb * (128 * 128), col0 + kb * 64, grow0);
    bulk_commit();
    if (has_add && has_next) {
      // the other staging tile was last read by the store of tile it-1: all but the newest bulk group (this tile's store) done
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      load_addend(out_base + (buf ^ 1) * kOutBytes, res_bar0 + 8u * (buf ^ 1), next_col0, next_grow0);
    }
  }
}

// Can this GEMM use the staged epilogue?  Returns the addend mode (-1: no) and the addend tile source.
inline int staged_epilogue_mode(const GemmDesc& g, const __nv_bfloat16** add, int* add_rows, int* add_ld) {
  *add = nullptr; *add_rows = 0; *add_ld = 0;
  const bool bf16_out = g.epi == EPI_STORE_BF16 || g.epi == EPI_BIAS_GELU || g.epi == EPI_MUL_GELU_GRAD;
  const bool aligned = (g.ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && (!g.C2 || (reinterpret_cast<uintptr_t>(g.C2) & 15) == 0);
  if (!bf16_out || !aligned || (g.pos && g.residual)) return -1;
  const bool res_ok = g.residual && (g.ldr % 8) == 0 && (reinterpret_cast<uintptr_t>(g.residual) & 15) == 0;
  if (g.epi == EPI_MUL_GELU_GRAD) {
    if (!res_ok || g.pos) return -1;
    *add = g.residual; *add_rows = g.M; *add_ld = g.ldr;
    return 3;
  }
  if (g.residual) {
    if (!res_ok || g.epi != EPI_STORE_BF16) return -1;
    *add = g.residual; *add_rows = g.M; *add_ld = g.ldr;
    return 1;
  }
  if (g.pos) {
    if (!g.pos_tile || g.epi != EPI_STORE_BF16) return -1;
    *add = g.pos_tile; *add_rows = 128; *add_ld = g.N;
    return 2;
  }
  return 0;
}

}  // namespace vitmarl
