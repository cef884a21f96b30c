// Bandwidth-bound helpers of the ViT encoder: patchify / unpatchify, LayerNorm forward and
// backward (warp per token row, fp32 statistics), final LayerNorm + mean-pool, bias-gradient
// column sums, parameter casts.  All activations are bf16; statistics / parameters fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {


// ---------------------------------------------------------------- patchify
// out[(b*T + py*Wp + px), (ph*P + pw)*C + c] = x[b, py*P+ph, px*P+pw, c]; 16-byte chunks (P*C % 8 == 0)
__global__ void patchify_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int C, int P, bool inverse) {
  const int chunks_per_prow = P * C / 8;              // chunks in one patch row (contiguous in the image)
  const int Wp = W / P, Hp = H / P;
  const size_t total = (size_t)B * Hp * Wp * P * chunks_per_prow;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int ck = r % chunks_per_prow; r /= chunks_per_prow;
    const int ph = r % P; r /= P;
    const int px = r % Wp; r /= Wp;
    const int py = r % Hp; r /= Hp;
    const int b = (int)r;
    const size_t img = (((size_t)b * H + py * P + ph) * W + px * P) * C / 8 + ck;
    if (!inverse) out[i] = __ldg(x + img);
    else out[img] = __ldg(x + i);
  }
}

int launch_patchify(cudaStream_t s, const __nv_bfloat16* x, __nv_bfloat16* out, int B, int H, int W, int C, int P) {
  if ((P * C) % 8 || H % P || W % P) { set_last_error("patchify: P*C must be a multiple of 8 and P must divide H, W"); return VITMARL_EINVAL; }
  const size_t total = (size_t)B * H * W * C / 8;
  if (!total) return VITMARL_OK;
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16);
  patchify_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), B, H, W, C, P, false);
  return check_cuda(cudaGetLastError());
}
int launch_unpatchify(cudaStream_t s, const __nv_bfloat16* dp, __nv_bfloat16* dx, int B, int H, int W, int C, int P) {
  if ((P * C) % 8 || H % P || W % P) return VITMARL_EINVAL;
  const size_t total = (size_t)B * H * W * C / 8;
  if (!total) return VITMARL_OK;
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16);
  patchify_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(dp), reinterpret_cast<uint4*>(dx), B, H, W, C, P, true);
  return check_cuda(cudaGetLastError());
}

// ---------------------------------------------------------------- LayerNorm helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NP>
__device__ __forceinline__ void load_row(const __nv_bfloat16* row, int lane, float (&v)[2 * NP]) {
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(row) + lane + 32 * i);
    v[2 * i] = bf16_lo(w);
    v[2 * i + 1] = bf16_hi(w);
  }
}

template <int NP>
__device__ __forceinline__ void row_stats(const float (&v)[2 * NP], int D, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * NP; ++i) s += v[i];
  mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * NP; ++i) { float d = v[i] - mean; q += d * d; }
  rstd = rsqrtf(warp_sum(q) / D + eps);
}

template <int NP>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                         float* __restrict__ stats, int M, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float g[2 * NP], bt[2 * NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float2 a = __ldg(reinterpret_cast<const float2*>(gamma) + lane + 32 * i), c = __ldg(reinterpret_cast<const float2*>(beta) + lane + 32 * i);
    g[2 * i] = a.x; g[2 * i + 1] = a.y; bt[2 * i] = c.x; bt[2 * i + 1] = c.y;
  }
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    float v[2 * NP];
    load_row<NP>(x + (size_t)row * D, lane, v);
    float mean, rstd;
    row_stats<NP>(v, D, eps, mean, rstd);
    uint32_t* out = reinterpret_cast<uint32_t*>(y + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NP; ++i)
      out[lane + 32 * i] = pack_bf16((v[2 * i] - mean) * rstd * g[2 * i] + bt[2 * i], (v[2 * i + 1] - mean) * rstd * g[2 * i + 1] + bt[2 * i + 1]);
    if (stats && lane == 0) reinterpret_cast<float2*>(stats)[row] = make_float2(mean, rstd);
  }
}

// dx = dx_add + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  dgamma += dy * xhat, dbeta += dy
template <int NP>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ stats, const __nv_bfloat16* __restrict__ dy,
                                                             const __nv_bfloat16* __restrict__ dx_add, __nv_bfloat16* __restrict__ dx,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int D) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float g[2 * NP], dg[2 * NP], db[2 * NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float2 a = __ldg(reinterpret_cast<const float2*>(gamma) + lane + 32 * i);
    g[2 * i] = a.x; g[2 * i + 1] = a.y;
    dg[2 * i] = dg[2 * i + 1] = db[2 * i] = db[2 * i + 1] = 0.f;
  }
  // The three row streams (x, dy, residual gradient) of the NEXT row are requested before the current row is reduced:
  // with ~100 registers only 16 warps fit an SM, so each warp keeps two rows (36 x 128 B) in flight.
  const int row0 = blockIdx.x * wpb + (threadIdx.x >> 5), stride = gridDim.x * wpb;
  uint32_t nx[NP], nd[NP], na[NP];
  auto fetch = [&](int row) {
    if (row < M) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        nx[i] = __ldg(reinterpret_cast<const uint32_t*>(x + (size_t)row * D) + lane + 32 * i);
        nd[i] = __ldg(reinterpret_cast<const uint32_t*>(dy + (size_t)row * D) + lane + 32 * i);
        na[i] = dx_add ? __ldg(reinterpret_cast<const uint32_t*>(dx_add + (size_t)row * D) + lane + 32 * i) : 0u;
      }
    }
  };
  fetch(row0);
  for (int row = row0; row < M; row += stride) {
    float v[2 * NP], d[2 * NP];
    uint32_t a[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      v[2 * i] = bf16_lo(nx[i]); v[2 * i + 1] = bf16_hi(nx[i]);
      d[2 * i] = bf16_lo(nd[i]); d[2 * i + 1] = bf16_hi(nd[i]);
      a[i] = na[i];
    }
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + row);
    fetch(row + stride);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) {
      v[i] = (v[i] - st.x) * st.y;           // xhat
      dg[i] += d[i] * v[i];
      db[i] += d[i];
      d[i] *= g[i];                          // g
      s1 += d[i];
      s2 += d[i] * v[i];
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    uint32_t* out = reinterpret_cast<uint32_t*>(dx + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float r0 = st.y * (d[2 * i] - s1 - v[2 * i] * s2) + bf16_lo(a[i]), r1 = st.y * (d[2 * i + 1] - s1 - v[2 * i + 1] * s2) + bf16_hi(a[i]);
      out[lane + 32 * i] = pack_bf16(r0, r1);
    }
  }
  // block-level reduction of the column partials, then one atomic per column per CTA
  extern __shared__ float red[];   // [2][D]
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c = 2 * (lane + 32 * i);
    atomicAdd(&red[c], dg[2 * i]); atomicAdd(&red[c + 1], dg[2 * i + 1]);
    atomicAdd(&red[D + c], db[2 * i]); atomicAdd(&red[D + c + 1], db[2 * i + 1]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[D + i]);
  }
}


// ---------------------------------------------------------------- LayerNorm, 16-byte-load variants for D = 24 * LANES * 8
// (D = 192: 8 lanes per row, D = 384: 16, D = 768: 32; each lane owns 3 chunks of 8 features: chunk = lane_in_group + LANES * i).
// The warp-per-row kernels above move 4 bytes per lane per load instruction and sit at ~67 % of the HBM roofline; these issue
// a quarter of the load instructions and keep 32 / LANES rows in flight per warp.
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = 1; o < LANES; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& w, float (&f)[8]) {
  f[0] = bf16_lo(w.x); f[1] = bf16_hi(w.x); f[2] = bf16_lo(w.y); f[3] = bf16_hi(w.y);
  f[4] = bf16_lo(w.z); f[5] = bf16_hi(w.z); f[6] = bf16_lo(w.w); f[7] = bf16_hi(w.w);
}

template <int LANES>
__global__ void __launch_bounds__(256, 2) layernorm16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                           float* __restrict__ stats, int M, float eps) {
  constexpr int D = LANES * 24, RPW = 32 / LANES;                    // rows per warp pass
  const int lane = threadIdx.x & 31, lg = lane % LANES, grp = lane / LANES;
  float g[3][8], bt[3][8];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { g[i][e] = gamma[(lg + LANES * i) * 8 + e]; bt[i][e] = beta[(lg + LANES * i) * 8 + e]; }
  const int wpb = blockDim.x >> 5;
  for (int row = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RPW + grp; row < M; row += gridDim.x * wpb * RPW) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * D);
    float v[3][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4 w = __ldg(xr + lg + LANES * i);
      unpack8(w, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
    const float mean = group_sum<LANES>(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; q += d * d; }
    const float rstd = rsqrtf(group_sum<LANES>(q) * (1.0f / D) + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * g[i][e] + bt[i][e];
      yr[lg + LANES * i] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
    if (stats && lg == 0) reinterpret_cast<float2*>(stats)[row] = make_float2(mean, rstd);
  }
}

template <int LANES>
__global__ void __launch_bounds__(256, 2) layernorm16_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ stats, const __nv_bfloat16* __restrict__ dy,
                                                               const __nv_bfloat16* __restrict__ dx_add, __nv_bfloat16* __restrict__ dx,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta, int M) {
  constexpr int D = LANES * 24, RPW = 32 / LANES;
  const int lane = threadIdx.x & 31, lg = lane % LANES, grp = lane / LANES;
  float g[3][8], dg[3][8], db[3][8];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { g[i][e] = gamma[(lg + LANES * i) * 8 + e]; dg[i][e] = 0.f; db[i][e] = 0.f; }
  const int wpb = blockDim.x >> 5;
  for (int row = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RPW + grp; row < M; row += gridDim.x * wpb * RPW) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * D);
    const uint4* yr = reinterpret_cast<const uint4*>(dy + (size_t)row * D);
    uint4 wx[3], wy[3], wa[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      wx[i] = __ldg(xr + lg + LANES * i);
      wy[i] = __ldg(yr + lg + LANES * i);
      wa[i] = dx_add ? __ldg(reinterpret_cast<const uint4*>(dx_add + (size_t)row * D) + lg + LANES * i) : make_uint4(0u, 0u, 0u, 0u);
    }
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + row);
    float v[3][8], d[3][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      unpack8(wx[i], v[i]);
      unpack8(wy[i], d[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[i][e] = (v[i][e] - st.x) * st.y;     // xhat
        dg[i][e] += d[i][e] * v[i][e];
        db[i][e] += d[i][e];
        d[i][e] *= g[i][e];                    // g
        s1 += d[i][e];
        s2 += d[i][e] * v[i][e];
      }
    }
    s1 = group_sum<LANES>(s1) * (1.0f / D);
    s2 = group_sum<LANES>(s2) * (1.0f / D);
    uint4* outr = reinterpret_cast<uint4*>(dx + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float a[8], o[8];
      unpack8(wa[i], a);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = st.y * (d[i][e] - s1 - v[i][e] * s2) + a[e];
      outr[lg + LANES * i] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
  // block-level reduction of the column partials, then one atomic per column per CTA
  extern __shared__ float red[];   // [2][D]
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (lg + LANES * i) * 8 + e;
      atomicAdd(&red[c], dg[i][e]);
      atomicAdd(&red[D + c], db[i][e]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[D + i]);
  }
}

#define VM_DISPATCH_NP(D, CALL)                                      \
  switch ((D) / 64) {                                                \
    case 1: { constexpr int NP = 1; CALL; } break;                   \
    case 2: { constexpr int NP = 2; CALL; } break;                   \
    case 3: { constexpr int NP = 3; CALL; } break;                   \
    case 4: { constexpr int NP = 4; CALL; } break;                   \
    case 6: { constexpr int NP = 6; CALL; } break;                   \
    case 8: { constexpr int NP = 8; CALL; } break;                   \
    case 12: { constexpr int NP = 12; CALL; } break;                 \
    default: set_last_error("layernorm: D/64 must be in {1,2,3,4,6,8,12}"); return VITMARL_EINVAL; \
  }

int launch_layernorm(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y, float* stats,
                     int M, int D, float eps) {
  if (M <= 0) return VITMARL_OK;
  if (D % 64) { set_last_error("layernorm: D % 64 != 0"); return VITMARL_EINVAL; }
  const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (al16 && (D == 192 || D == 384 || D == 768)) {
    const int lanes = D / 24, rows_per_cta = 8 * (32 / lanes);
    const int grid16 = min((M + rows_per_cta - 1) / rows_per_cta, num_sms() * 8);
    if (lanes == 8) layernorm16_kernel<8><<<grid16, 256, 0, s>>>(x, gamma, beta, y, stats, M, eps);
    else if (lanes == 16) layernorm16_kernel<16><<<grid16, 256, 0, s>>>(x, gamma, beta, y, stats, M, eps);
    else layernorm16_kernel<32><<<grid16, 256, 0, s>>>(x, gamma, beta, y, stats, M, eps);
    return check_cuda(cudaGetLastError());
  }
  const int grid = min((M + 7) / 8, num_sms() * 8);
  VM_DISPATCH_NP(D, (layernorm_kernel<NP><<<grid, 256, 0, s>>>(x, gamma, beta, y, stats, M, D, eps)));
  return check_cuda(cudaGetLastError());
}

int launch_layernorm_bwd(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* stats, const __nv_bfloat16* dy,
                         const __nv_bfloat16* dx_add, __nv_bfloat16* dx, float* dgamma, float* dbeta, int M, int D) {
  if (M <= 0) return VITMARL_OK;
  if (D % 64) return VITMARL_EINVAL;
  const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                      reinterpret_cast<uintptr_t>(dx_add)) & 15) == 0;
  if (al16 && (D == 192 || D == 384 || D == 768)) {
    const int lanes = D / 24, rows_per_cta = 8 * (32 / lanes);
    const int grid16 = min((M + rows_per_cta - 1) / rows_per_cta, num_sms() * 4);
    const size_t sm = 2 * D * sizeof(float);
    if (lanes == 8) layernorm16_bwd_kernel<8><<<grid16, 256, sm, s>>>(x, gamma, stats, dy, dx_add, dx, dgamma, dbeta, M);
    else if (lanes == 16) layernorm16_bwd_kernel<16><<<grid16, 256, sm, s>>>(x, gamma, stats, dy, dx_add, dx, dgamma, dbeta, M);
    else layernorm16_bwd_kernel<32><<<grid16, 256, sm, s>>>(x, gamma, stats, dy, dx_add, dx, dgamma, dbeta, M);
    return check_cuda(cudaGetLastError());
  }
  const int grid = min((M + 7) / 8, num_sms() * 4);
  VM_DISPATCH_NP(D, (layernorm_bwd_kernel<NP><<<grid, 256, 2 * D * sizeof(float), s>>>(x, gamma, stats, dy, dx_add, dx, dgamma, dbeta, M, D)));
  return check_cuda(cudaGetLastError());
}

// ---------------------------------------------------------------- final LN + mean pool (one CTA of 8 warps per image)
// EIGHT lanes per token row (each lane D/8 features as D/64 16-byte chunks), 32 rows in flight per CTA: a row of 192
// features is too little work for a whole warp -- with one row per warp the kernel was bound by the two 5-step shuffle
// reductions per row (45 us for 100 MB), not by memory.  Per-group partial sums are combined in a fixed order
// (bitwise deterministic).
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <int NP>
__global__ void __launch_bounds__(256, NP <= 4 ? 2 : 1) final_ln_pool_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float* __restrict__ y,
                                                                float* __restrict__ stats, int B, int T, int D, float eps) {
  extern __shared__ float red[];   // [32 row groups][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, grp = warp * 4 + (lane >> 3);
  // Persistent CTAs (thousands of 2-us CTAs were mostly launch and tail): the rows of the NEXT image are requested before
  // the current one is reduced.  Up to 64 tokens (2 rows per 8-lane group) are held in registers at a time.
  constexpr int RPG = NP <= 3 ? 2 : 1;
  uint4 nxt[RPG][NP];
  auto fetch = [&](int b, int t0) {
    if (b < B) {
#pragma unroll
      for (int r = 0; r < RPG; ++r) {
        const int t = t0 + r * 32 + grp;
        const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)(b * T + (t < T ? t : 0)) * D);
#pragma unroll
        for (int i = 0; i < NP; ++i) nxt[r][i] = __ldg(src + sub + 8 * i);
      }
    }
  };
  fetch(blockIdx.x, 0);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float acc[8 * NP];
#pragma unroll
    for (int i = 0; i < 8 * NP; ++i) acc[i] = 0.f;
    for (int t0 = 0; t0 < T; t0 += 32 * RPG) {           // uniform trip count: the group shuffles run under a full mask
      uint4 cur[RPG][NP];
#pragma unroll
      for (int r = 0; r < RPG; ++r)
#pragma unroll
        for (int i = 0; i < NP; ++i) cur[r][i] = nxt[r][i];
      if (t0 + 32 * RPG < T) fetch(b, t0 + 32 * RPG); else fetch(b + gridDim.x, 0);
#pragma unroll
      for (int r = 0; r < RPG; ++r) {
        const int t = t0 + r * 32 + grp;
        const bool valid = t < T;
        float v[8 * NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const uint4 w = cur[r][i];
          v[8 * i] = bf16_lo(w.x); v[8 * i + 1] = bf16_hi(w.x); v[8 * i + 2] = bf16_lo(w.y); v[8 * i + 3] = bf16_hi(w.y);
          v[8 * i + 4] = bf16_lo(w.z); v[8 * i + 5] = bf16_hi(w.z); v[8 * i + 6] = bf16_lo(w.w); v[8 * i + 7] = bf16_hi(w.w);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8 * NP; ++i) s += v[i];
        const float mean = group8_sum(s) / D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8 * NP; ++i) { const float d = v[i] - mean; q += d * d; }
        const float rstd = rsqrtf(group8_sum(q) / D + eps);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 8 * NP; ++i) acc[i] += (v[i] - mean) * rstd;
          if (stats && sub == 0) reinterpret_cast<float2*>(stats)[b * T + t] = make_float2(mean, rstd);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float4* dst = reinterpret_cast<float4*>(red + (size_t)grp * D + (sub + 8 * i) * 8);
      dst[0] = make_float4(acc[8 * i], acc[8 * i + 1], acc[8 * i + 2], acc[8 * i + 3]);
      dst[1] = make_float4(acc[8 * i + 4], acc[8 * i + 5], acc[8 * i + 6], acc[8 * i + 7]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      float s = 0.f;
#pragma unroll 8
      for (int g = 0; g < 32; ++g) s += red[g * D + i];
      y[(size_t)b * D + i] = s / T * __ldg(gamma + i) + __ldg(beta + i);
    }
    __syncthreads();                                     // red is rewritten for the next image
  }
}

// dx[t,:] = LN_bwd(dy[b,:] / T);  dgamma += sum_t dy/T * xhat,  dbeta += dy   (dbeta: sum over tokens of dy/T = dy)
template <int NP>
__global__ void __launch_bounds__(256) final_ln_pool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ stats, const float* __restrict__ dy,
                                                                 __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, int T, int D) {
  extern __shared__ float red[];   // [D]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float g[2 * NP], d0[2 * NP], dg[2 * NP];
  const float invT = 1.0f / T;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float2 a = __ldg(reinterpret_cast<const float2*>(gamma) + lane + 32 * i);
    float2 c = __ldg(reinterpret_cast<const float2*>(dy + (size_t)b * D) + lane + 32 * i);
    g[2 * i] = a.x; g[2 * i + 1] = a.y;
    d0[2 * i] = c.x * invT; d0[2 * i + 1] = c.y * invT;
    dg[2 * i] = dg[2 * i + 1] = 0.f;
  }
  for (int t = warp; t < T; t += nw) {
    const int row = b * T + t;
    float v[2 * NP], d[2 * NP];
    load_row<NP>(x + (size_t)row * D, lane, v);
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + row);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) {
      v[i] = (v[i] - st.x) * st.y;
      dg[i] += d0[i] * v[i];
      d[i] = d0[i] * g[i];
      s1 += d[i];
      s2 += d[i] * v[i];
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    uint32_t* out = reinterpret_cast<uint32_t*>(dx + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NP; ++i)
      out[lane + 32 * i] = pack_bf16(st.y * (d[2 * i] - s1 - v[2 * i] * s2), st.y * (d[2 * i + 1] - s1 - v[2 * i + 1] * s2));
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    atomicAdd(&red[2 * (lane + 32 * i)], dg[2 * i]);
    atomicAdd(&red[2 * (lane + 32 * i) + 1], dg[2 * i + 1]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, __ldg(dy + (size_t)b * D + i));
  }
}

template <int NP>
static void launch_final_ln_pool_t(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* beta, float* y, float* stats,
                                   int B, int T, int D, float eps, size_t smem) {
  if (smem > 48 * 1024 && cudaFuncSetAttribute(final_ln_pool_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;
  final_ln_pool_kernel<NP><<<min(B, 2 * num_sms()), 256, smem, s>>>(x, gamma, beta, y, stats, B, T, D, eps);
}

int launch_final_ln_pool(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* beta, float* y, float* stats,
                         int B, int T, int D, float eps) {
  if (B <= 0) return VITMARL_OK;
  if (D % 64) return VITMARL_EINVAL;
  const size_t smem = 32 * (size_t)D * sizeof(float);
  if ((reinterpret_cast<uintptr_t>(x) & 15)) return VITMARL_EINVAL;
  VM_DISPATCH_NP(D, (launch_final_ln_pool_t<NP>(s, x, gamma, beta, y, stats, B, T, D, eps, smem)));
  return check_cuda(cudaGetLastError());
}
int launch_final_ln_pool_bwd(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* stats, const float* dy,
                             __nv_bfloat16* dx, float* dgamma, float* dbeta, int B, int T, int D) {
  if (B <= 0) return VITMARL_OK;
  if (D % 64) return VITMARL_EINVAL;
  VM_DISPATCH_NP(D, (final_ln_pool_bwd_kernel<NP><<<B, 256, D * sizeof(float), s>>>(x, gamma, stats, dy, dx, dgamma, dbeta, T, D)));
  return check_cuda(cudaGetLastError());
}

// ---------------------------------------------------------------- elementwise / reductions
__global__ void gelu_bwd_kernel(const uint4* __restrict__ pre, const uint4* __restrict__ dy, uint4* __restrict__ dx, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4 p = __ldg(pre + i), d = __ldg(dy + i), o;
    const uint32_t* pp = &p.x; const uint32_t* dd = &d.x; uint32_t* oo = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      oo[k] = pack_bf16(bf16_lo(dd[k]) * gelu_tanh_grad(bf16_lo(pp[k])), bf16_hi(dd[k]) * gelu_tanh_grad(bf16_hi(pp[k])));
    dx[i] = o;
  }
}
int launch_gelu_bwd(cudaStream_t s, const __nv_bfloat16* pre, const __nv_bfloat16* dy, __nv_bfloat16* dx, size_t n) {
  if (!n) return VITMARL_OK;
  if (n % 8) return VITMARL_EINVAL;
  const int grid = (int)std::min<size_t>((n / 8 + 255) / 256, (size_t)num_sms() * 16);
  gelu_bwd_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(pre), reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), n / 8);
  return check_cuda(cudaGetLastError());
}

// out[n] += sum_m x[m, n]; each thread owns a bf16 pair of columns, CTAs stride over row blocks
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int M, int N, int rows_per_cta) {
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  for (int c = threadIdx.x; c < N / 2; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int r = r0; r < r1; ++r) {
      uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(x + (size_t)r * N) + c);
      a += bf16_lo(w); b += bf16_hi(w);
    }
    atomicAdd(out + 2 * c, a);
    atomicAdd(out + 2 * c + 1, b);
  }
}
int launch_colsum(cudaStream_t s, const __nv_bfloat16* x, float* out, int M, int N) {
  if (M <= 0) return VITMARL_OK;
  if (N % 2) return VITMARL_EINVAL;
  const int ctas = min(M, num_sms() * 8);
  const int rpc = (M + ctas - 1) / ctas;
  colsum_kernel<<<(M + rpc - 1) / rpc, 256, 0, s>>>(x, out, M, N, rpc);
  return check_cuda(cudaGetLastError());
}

__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
}
int launch_cast_f32_to_bf16(cudaStream_t s, const float* src, __nv_bfloat16* dst, size_t n) {
  if (!n) return VITMARL_OK;
  cast_kernel<<<(int)std::min<size_t>((n + 255) / 256, 4096), 256, 0, s>>>(src, dst, n);
  return check_cuda(cudaGetLastError());
}

__global__ void transpose_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)c * rows + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
int launch_transpose_f32_to_bf16(cudaStream_t s, const float* src, __nv_bfloat16* dst, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return VITMARL_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_cast_kernel<<<grid, block, 0, s>>>(src, dst, rows, cols);
  return check_cuda(cudaGetLastError());
}

}  // namespace vitmarl
