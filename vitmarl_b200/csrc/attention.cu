// Multi-head self-attention for the ViT encoder: T = 64 tokens per image, head dim 64.
// One CTA per (image, head); the whole problem lives in shared memory / registers.
// v0 uses warp-level mma.sync.m16n8k16 (bf16 -> fp32); ~5 % of the encoder FLOPs.
// (The fused tcgen05 block kernel replaces this for the forward path; see DESIGN.md.)
//
// forward : S = Q K^T / 8, P = softmax(S), O = P V
// backward: recompute S, P;  dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(dP o P));
//           dQ = dS K / 8;  dK = dS^T Q / 8
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

constexpr int AT = 64;        // tokens
constexpr int ADH = 64;       // head dim
constexpr int APITCH = 72;    // smem row pitch in bf16 (144 B): conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// cooperative 64x64 bf16 tile load global -> padded smem (128 threads, 16-byte chunks)
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int ld, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = tid + i * 128, r = c >> 3, cc = (c & 7) * 8;
    *reinterpret_cast<uint4*>(s + r * APITCH + cc) = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * ld + cc));
  }
}
__device__ __forceinline__ void store_rows16(__nv_bfloat16* g, int ld, const __nv_bfloat16* s, int row0, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = lane + i * 32, r = row0 + (c >> 3), cc = (c & 7) * 8;
    *reinterpret_cast<uint4*>(g + (size_t)r * ld + cc) = *reinterpret_cast<const uint4*>(s + r * APITCH + cc);
  }
}

// C[16 x 64] = A[16 rows of sA starting row0, 64 k] . B^T where B element (n, k) = sB[n][k]   (both k-contiguous)
__device__ __forceinline__ void mm_nt(float (&c)[8][4], const __nv_bfloat16* sA, int row0, const __nv_bfloat16* sB, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(smem_u32(sA + (row0 + (lane & 15)) * APITCH + ks * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
    for (int nt = 0; nt < 8; nt += 2) {
      uint32_t b0, b1, b2, b3;   // (nt: k lo, k hi), (nt+1: k lo, k hi)
      ldsm_x4(smem_u32(sB + (nt * 8 + (lane & 7) + ((lane >> 4) & 1) * 8) * APITCH + ks * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
      mma_bf16(c[nt], a0, a1, a2, a3, b0, b1);
      mma_bf16(c[nt + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// C[16 x 64] += A(regs, 16 x 64 as 4 k-tiles of A fragments) . B where B element (k, n) = sB[k][n]   (n-contiguous)
__device__ __forceinline__ void mm_nn_regA(float (&c)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16* sB, int lane) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
    for (int nt = 0; nt < 8; nt += 2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sB + (kt * 16 + (lane & 15)) * APITCH + (nt + (lane >> 4)) * 8), b0, b1, b2, b3);
      mma_bf16(c[nt], a[kt][0], a[kt][1], a[kt][2], a[kt][3], b0, b1);
      mma_bf16(c[nt + 1], a[kt][0], a[kt][1], a[kt][2], a[kt][3], b2, b3);
    }
  }
}

// C[16 x 64] = A^T . B with A element (k, m) = sA[k][m0 + m] (m-contiguous), B element (k, n) = sB[k][n]
__device__ __forceinline__ void mm_tn(float (&c)[8][4], const __nv_bfloat16* sA, int m0, const __nv_bfloat16* sB, int lane) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t a0, a1, a2, a3;
    // matrices: (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15), each transposed on load
    ldsm_x4_t(smem_u32(sA + (kt * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * APITCH + m0 + ((lane >> 3) & 1) * 8), a0, a1, a2, a3);
#pragma unroll
    for (int nt = 0; nt < 8; nt += 2) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(smem_u32(sB + (kt * 16 + (lane & 15)) * APITCH + (nt + (lane >> 4)) * 8), b0, b1, b2, b3);
      mma_bf16(c[nt], a0, a1, a2, a3, b0, b1);
      mma_bf16(c[nt + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
}

// row-wise softmax of the warp's 16 x 64 score tile held in mma C layout; returns P (normalised)
__device__ __forceinline__ void softmax_rows(float (&s)[8][4]) {
  const float sl2 = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) { m0 = fmaxf(m0, fmaxf(s[i][0], s[i][1])); m1 = fmaxf(m1, fmaxf(s[i][2], s[i][3])); }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s[i][0] = exp2f((s[i][0] - m0) * sl2); s[i][1] = exp2f((s[i][1] - m0) * sl2);
    s[i][2] = exp2f((s[i][2] - m1) * sl2); s[i][3] = exp2f((s[i][3] - m1) * sl2);
    l0 += s[i][0] + s[i][1]; l1 += s[i][2] + s[i][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float r0 = 1.f / l0, r1 = 1.f / l1;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i][0] *= r0; s[i][1] *= r0; s[i][2] *= r1; s[i][3] *= r1; }
}

// mma C layout (16 x 64 fp32) -> A fragments (bf16) for a following 16 x 64 x N product
__device__ __forceinline__ void c_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    a[kt][0] = pack_bf16(c[2 * kt][0], c[2 * kt][1]);
    a[kt][1] = pack_bf16(c[2 * kt][2], c[2 * kt][3]);
    a[kt][2] = pack_bf16(c[2 * kt + 1][0], c[2 * kt + 1][1]);
    a[kt][3] = pack_bf16(c[2 * kt + 1][2], c[2 * kt + 1][3]);
  }
}

// mma C layout -> bf16 rows in padded smem (rows row0 .. row0+15)
__device__ __forceinline__ void c_to_smem(const float (&c)[8][4], __nv_bfloat16* s, int row0, int lane, float scale) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(s + (row0 + g) * APITCH + nt * 8 + 2 * t) = pack_bf16(c[nt][0] * scale, c[nt][1] * scale);
    *reinterpret_cast<uint32_t*>(s + (row0 + g + 8) * APITCH + nt * 8 + 2 * t) = pack_bf16(c[nt][2] * scale, c[nt][3] * scale);
  }
}

__global__ void __launch_bounds__(128) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                        int heads, int D) {
  __shared__ __align__(16) __nv_bfloat16 sQ[AT * APITCH], sK[AT * APITCH], sV[AT * APITCH];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld = 3 * D;
  const __nv_bfloat16* base = qkv + (size_t)b * AT * ld + h * ADH;
  load_tile(sQ, base, ld, tid);
  load_tile(sK, base + D, ld, tid);
  load_tile(sV, base + 2 * D, ld, tid);
  __syncthreads();
  float s[8][4];
  zero_acc(s);
  mm_nt(s, sQ, warp * 16, sK, lane);
  softmax_rows(s);
  uint32_t pa[4][4];
  c_to_a(s, pa);
  float o[8][4];
  zero_acc(o);
  mm_nn_regA(o, pa, sV, lane);
  __syncwarp();
  c_to_smem(o, sQ, warp * 16, lane, 1.0f);     // this warp's Q rows are no longer needed
  __syncwarp();
  store_rows16(out + (size_t)b * AT * D + h * ADH, D, sQ, warp * 16, lane);
}

__global__ void __launch_bounds__(128) attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                        __nv_bfloat16* __restrict__ dqkv, int heads, int D) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_dyn);
  __nv_bfloat16* sK = sQ + AT * APITCH;
  __nv_bfloat16* sV = sK + AT * APITCH;
  __nv_bfloat16* sdO = sV + AT * APITCH;
  __nv_bfloat16* sP = sdO + AT * APITCH;
  __nv_bfloat16* sdS = sP + AT * APITCH;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld = 3 * D, row0 = warp * 16;
  const __nv_bfloat16* base = qkv + (size_t)b * AT * ld + h * ADH;
  load_tile(sQ, base, ld, tid);
  load_tile(sK, base + D, ld, tid);
  load_tile(sV, base + 2 * D, ld, tid);
  load_tile(sdO, dout + (size_t)b * AT * D + h * ADH, D, tid);
  __syncthreads();
  // P (16 x 64 per warp)
  float p[8][4];
  zero_acc(p);
  mm_nt(p, sQ, row0, sK, lane);
  softmax_rows(p);
  // dP = dO V^T
  float dp[8][4];
  zero_acc(dp);
  mm_nt(dp, sdO, row0, sV, lane);
  // D = rowsum(dP o P);  dS = P o (dP - D)
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { d0 += dp[i][0] * p[i][0] + dp[i][1] * p[i][1]; d1 += dp[i][2] * p[i][2] + dp[i][3] * p[i][3]; }
  d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dp[i][0] = p[i][0] * (dp[i][0] - d0); dp[i][1] = p[i][1] * (dp[i][1] - d0);
    dp[i][2] = p[i][2] * (dp[i][2] - d1); dp[i][3] = p[i][3] * (dp[i][3] - d1);
  }
  c_to_smem(p, sP, row0, lane, 1.0f);
  c_to_smem(dp, sdS, row0, lane, 1.0f);
  // dQ = dS K / 8   (A from registers)
  uint32_t a[4][4];
  c_to_a(dp, a);
  float acc[8][4];
  zero_acc(acc);
  mm_nn_regA(acc, a, sK, lane);
  __syncthreads();                       // all warps: P, dS complete; everyone has finished reading sQ rows of S = Q K^T
  // dV = P^T dO (this warp: keys row0 .. row0+15)
  float dv[8][4];
  zero_acc(dv);
  mm_tn(dv, sP, row0, sdO, lane);
  // dK = dS^T Q / 8
  float dk[8][4];
  zero_acc(dk);
  mm_tn(dk, sdS, row0, sQ, lane);
  __syncthreads();                       // sQ / sP / sdS no longer read -> reuse as output staging
  c_to_smem(acc, sQ, row0, lane, 0.125f);
  c_to_smem(dk, sP, row0, lane, 0.125f);
  c_to_smem(dv, sdS, row0, lane, 1.0f);
  __syncwarp();
  __nv_bfloat16* obase = dqkv + (size_t)b * AT * ld + h * ADH;
  store_rows16(obase, ld, sQ, row0, lane);
  store_rows16(obase + D, ld, sP, row0, lane);
  store_rows16(obase + 2 * D, ld, sdS, row0, lane);
}

int launch_attention(cudaStream_t s, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int heads) {
  if (B <= 0) return VITMARL_OK;
  attn_fwd_kernel<<<B * heads, 128, 0, s>>>(qkv, out, heads, heads * ADH);
  return check_cuda(cudaGetLastError());
}

int launch_attention_bwd(cudaStream_t s, const __nv_bfloat16* qkv, const __nv_bfloat16* dout, __nv_bfloat16* dqkv, int B, int heads) {
  if (B <= 0) return VITMARL_OK;
  const int smem = 6 * AT * APITCH * 2;
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return check_cuda(e);
  attn_bwd_kernel<<<B * heads, 128, smem, s>>>(qkv, dout, dqkv, heads, heads * ADH);
  return check_cuda(cudaGetLastError());
}

}  // namespace vitmarl
