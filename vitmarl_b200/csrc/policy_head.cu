// Policy head behind the encoder: the Dense / GRU pieces of
// `ActorCriticRNN.__call__(hidden, (obs, dones))` (gymnax_exchange/jaxrl/MARL/ippo_rnn_JAXMARL.py:48-115) in fp32 on the CUDA
// cores, so that the flax-shaped `apply(params, hidden, (obs, dones)) -> (hidden, logits, value)` of
// vitmarl_b200/actor_critic.py can route the rendered LOB image through the ViT encoder and feed its output into the first
// Dense without leaving the GPU or calling a framework GEMM.
//
// The head is ~0.35 MFLOP per (env, agent) row against 720 MFLOP for the row's ViT-Tiny encode (0.05 %), i.e. launch-bound
// small-matrix work: plain register-tiled fp32 kernels (64 x 64 CTA tiles, K staged through shared memory in steps of 16,
// 4 x 4 outputs per thread), bit-reproducible (fixed summation order), no tensor cores -- parity with the fp32 oracle is 1e-5.
//   vitmarl_dense_f32     y[R,N] = act( [x0 | x1][R,K0+K1] . W[K0+K1,N] + b )      (flax Dense kernel layout [in, out]; the
//                         two-source form is the concat(vector obs, encoding) input of the first Dense)
//   vitmarl_gru_cell_f32  flax.linen.GRUCell with the ScannedRNN reset: h <- where(reset, 0, h);
//                         r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r * (gh_n + b_hn)), h' = (1-z) n + z h
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

constexpr int kTile = 64, kStep = 16;

__global__ void __launch_bounds__(256) dense_f32_kernel(int R, int K0, int K1, int N, const float* __restrict__ x0, int ldx0,
                                                        const float* __restrict__ x1, int ldx1, const float* __restrict__ W,
                                                        const float* __restrict__ bias, int act, float* __restrict__ y, int ldy) {
  __shared__ float sx[kStep][kTile + 1];   // [k][row]
  __shared__ float sw[kStep][kTile];       // [k][col]
  const int K = K0 + K1;
  const int r0 = blockIdx.y * kTile, c0 = blockIdx.x * kTile;
  const int tr = (threadIdx.x >> 4) * 4, tc = (threadIdx.x & 15) * 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kStep) {
    for (int i = threadIdx.x; i < kTile * kStep; i += 256) {
      const int rr = i / kStep, kk = i % kStep, r = r0 + rr, k = k0 + kk;
      float v = 0.f;
      if (r < R && k < K) v = k < K0 ? x0[(size_t)r * ldx0 + k] : x1[(size_t)r * ldx1 + (k - K0)];
      sx[kk][rr] = v;
    }
    for (int i = threadIdx.x; i < kTile * kStep; i += 256) {
      const int kk = i / kTile, cc = i % kTile, k = k0 + kk, c = c0 + cc;
      sw[kk][cc] = (k < K && c < N) ? W[(size_t)k * N + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kStep; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sx[kk][tr + i]; b[i] = sw[kk][tc + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + tr + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tc + j;
      if (c >= N) continue;
      float v = acc[i][j] + (bias ? bias[c] : 0.f);
      if (act == VITMARL_ACT_RELU) v = fmaxf(v, 0.f);
      y[(size_t)r * ldy + c] = v;
    }
  }
}

__global__ void __launch_bounds__(256) gru_cell_kernel(int R, int H, const float* __restrict__ gi, const float* __restrict__ gh,
                                                       const float* __restrict__ b_hn, const float* __restrict__ h,
                                                       const uint8_t* __restrict__ reset, float* __restrict__ h_out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)R * H) return;
  const int r = (int)(i / H), c = (int)(i % H);
  const float* a = gi + (size_t)r * 3 * H;
  const float* b = gh + (size_t)r * 3 * H;
  const bool rs = reset && reset[r];
  // gh was computed from the carry AFTER the reset (a reset row's carry is zero, so its hidden products are zero)
  const float ghr = rs ? 0.f : b[c], ghz = rs ? 0.f : b[H + c], ghn = rs ? 0.f : b[2 * H + c];
  const float hp = rs ? 0.f : h[i];
  const float rg = 1.f / (1.f + expf(-(a[c] + ghr)));
  const float zg = 1.f / (1.f + expf(-(a[H + c] + ghz)));
  const float ng = tanhf(a[2 * H + c] + rg * (ghn + b_hn[c]));
  h_out[i] = (1.f - zg) * ng + zg * hp;
}

// Generalised advantage estimation of the PPO trainers (`_calculate_gae`, ippo_rnn_JAXMARL.py:372-394: a reverse lax.scan over
// NUM_STEPS): one thread per (env, agent) column walks the trajectory backwards.  Separate IEEE mul / add in the reference's
// order of operations (no FMA contraction), so the result is bit-identical to the NumPy restatement of the scan.
//     delta = reward + gamma * next_value * (1 - done) - value;   gae = delta + gamma * lambda * (1 - done) * gae
__global__ void __launch_bounds__(256) gae_kernel(int S, int B, float gamma, float lambda, const float* __restrict__ reward,
                                                  const float* __restrict__ value, const uint8_t* __restrict__ done,
                                                  const float* __restrict__ last_val, float* __restrict__ adv, float* __restrict__ target) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float gae = 0.f, next_value = last_val[b];
  const float gl = __fmul_rn(gamma, lambda);
  for (int t = S - 1; t >= 0; --t) {
    const size_t i = (size_t)t * B + b;
    const float nd = done[i] ? 0.f : 1.f, v = value[i];
    const float delta = __fsub_rn(__fadd_rn(reward[i], __fmul_rn(__fmul_rn(gamma, next_value), nd)), v);
    gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nd), gae));
    adv[i] = gae;
    if (target) target[i] = __fadd_rn(gae, v);
    next_value = v;
  }
}

}  // namespace vitmarl

using namespace vitmarl;

extern "C" int vitmarl_gae_f32(void* stream, int S, int B, float gamma, float gae_lambda, const float* reward, const float* value,
                               const uint8_t* done, const float* last_val, float* advantages, float* targets) {
  if (S == 0 || B == 0) return VITMARL_OK;
  if (S < 0 || B < 0 || !reward || !value || !done || !last_val || !advantages) return VITMARL_EINVAL;
  gae_kernel<<<(B + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, B, gamma, gae_lambda, reward, value, done, last_val, advantages, targets);
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_dense_f32(void* stream, int R, int K0, int K1, int N, const float* x0, int ldx0, const float* x1, int ldx1,
                                 const float* W, const float* bias, int act, float* y, int ldy) {
  if (R == 0) return VITMARL_OK;
  if (R < 0 || K0 < 1 || K1 < 0 || N < 1 || !x0 || (K1 > 0 && !x1) || !W || !y || ldx0 < K0 || (K1 > 0 && ldx1 < K1) || ldy < N ||
      (act != VITMARL_ACT_NONE && act != VITMARL_ACT_RELU))
    return VITMARL_EINVAL;
  const dim3 grid((N + kTile - 1) / kTile, (R + kTile - 1) / kTile);
  dense_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(R, K0, K1, N, x0, ldx0, x1, ldx1, W, bias, act, y, ldy);
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_gru_cell_f32(void* stream, int R, int H, const float* gi, const float* gh, const float* b_hn, const float* h,
                                    const uint8_t* reset, float* h_out) {
  if (R == 0) return VITMARL_OK;
  if (R < 0 || H < 1 || !gi || !gh || !b_hn || !h || !h_out) return VITMARL_EINVAL;
  const size_t n = (size_t)R * H;
  gru_cell_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(R, H, gi, gh, b_hn, h, reset, h_out);
  return check_cuda(cudaGetLastError());
}
