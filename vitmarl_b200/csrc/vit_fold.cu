// On-the-fly parameter folding for the fused inference kernels (one launch per forward, ~10 MB of traffic):
//     Wf[n, k]  = bf16( scale * W[n, k] * gamma[k] )            (LayerNorm scale folded into the following projection)
//     bf[n]     = scale * (bias[n] + sum_k W[n, k] * beta[k])   (LayerNorm shift folded into its bias)
// so that  (LN(x) * gamma + beta) . W^T + bias  ==  ((x - mean) * rstd) . Wf^T + bf   and the kernels' LayerNorm is a
// pure normalisation.  `scale` = 1/2 folds the 0.5 of gelu_tanh into FC2; `scale` = log2(e)/8 on the Q rows folds the softmax
// scale; a job with Wf = null only produces a bias (output bias + Wo . folded V bias: the V bias commutes with the
// row-stochastic P).  The caller's parameter table is never
// modified (the C ABI treats it as const; an optimiser may have updated it since the previous call).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

__global__ void __launch_bounds__(256) fold_params_kernel(const __grid_constant__ FoldJobs jobs) {
  const FoldJob& J = jobs.job[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= J.N) return;
  const __nv_bfloat162* w = reinterpret_cast<const __nv_bfloat162*>(J.W + (size_t)n * J.K);
  __nv_bfloat162* wf = reinterpret_cast<__nv_bfloat162*>(J.Wf + (size_t)n * J.K);
  float acc = 0.f;
  for (int k2 = lane; k2 < J.K / 2; k2 += 32) {
    const float2 v = __bfloat1622float2(w[k2]);
    const float g0 = J.gamma ? J.gamma[2 * k2] : 1.f, g1 = J.gamma ? J.gamma[2 * k2 + 1] : 1.f;
    if (J.Wf) wf[k2] = __floats2bfloat162_rn(J.scale * v.x * g0, J.scale * v.y * g1);
    if (J.beta) acc += v.x * J.beta[2 * k2] + v.y * J.beta[2 * k2 + 1];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0 && J.bias_out) {
    const float b = J.scale * ((J.bias ? J.bias[n] : 0.f) + acc);
    if (J.bias_out_bf16) reinterpret_cast<__nv_bfloat16*>(J.bias_out)[n] = __float2bfloat16_rn(b);
    else reinterpret_cast<float*>(J.bias_out)[n] = b;
  }
}

// pos [T, D] fp32 -> bf16 [128, D], rows repeated with period T (T divides 128): the addend tile of the patch-embedding GEMM
__global__ void __launch_bounds__(256) pos_tile_kernel(const float* __restrict__ pos, __nv_bfloat16* __restrict__ out, int T, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * D) return;
  const int r = i / D, c = i - r * D;
  out[i] = __float2bfloat16_rn(pos[(size_t)(r % T) * D + c]);
}

int launch_pos_tile(cudaStream_t s, const float* pos, __nv_bfloat16* out, int T, int D) {
  pos_tile_kernel<<<(128 * D + 255) / 256, 256, 0, s>>>(pos, out, T, D);
  return check_cuda(cudaGetLastError());
}

int launch_fold_params(cudaStream_t s, const FoldJobs& jobs) {
  if (jobs.n <= 0) return VITMARL_OK;
  int max_n = 0;
  for (int i = 0; i < jobs.n; ++i) max_n = jobs.job[i].N > max_n ? jobs.job[i].N : max_n;
  fold_params_kernel<<<dim3((max_n + 7) / 8, jobs.n), 256, 0, s>>>(jobs);
  return check_cuda(cudaGetLastError());
}

}  // namespace vitmarl
