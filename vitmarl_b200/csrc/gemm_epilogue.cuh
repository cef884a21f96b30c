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
  int add_mode;           // shared-memory-staged epilogue (gemm2): 0 no addend, 1 residual tile, 2 position-embedding tile
};


__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, uint32_t (&r)[32], int row, bool row_ok, int col) {
  if (p.epi == EPI_ATOMIC_F32) {
    if (row_ok) {
      float* dst = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col;
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(r[j]) * p.out_scale);
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

}  // namespace vitmarl
