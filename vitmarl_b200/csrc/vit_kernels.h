// Internal launch API shared by the ViT translation units (not part of the C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vitmarl {

enum GemmEpi {
  EPI_STORE_BF16 = 0,   // C(bf16) = acc (+bias) (+pos) (+residual)
  EPI_BIAS_GELU = 1,    // C(bf16) = gelu_tanh(acc + bias)
  EPI_STORE_F32 = 2,    // C(fp32) = acc (+bias)
  EPI_ATOMIC_F32 = 3,   // C(fp32) += out_scale * acc   (split-K over the contraction, red.add)
  EPI_MUL_GELU_GRAD = 4,// C(bf16) = acc * gelu_tanh'(aux[row,col]),  aux passed as `residual`
};

// C[M,N] = epi(A . B^T).  K-major operand X: element (row r, k) at X[r*ldx + k].
// MN-major operand X: element (row r, k) at X[k*ldx + r]  (i.e. stored as [K, rows]).
struct GemmDesc {
  int M = 0, N = 0, K = 0;
  const __nv_bfloat16* A = nullptr; int lda = 0; bool a_mn_major = false;
  const __nv_bfloat16* B = nullptr; int ldb = 0; bool b_mn_major = false;
  void* C = nullptr; int ldc = 0;
  void* C2 = nullptr;   // EPI_BIAS_GELU: also store the pre-activation (bf16) here
  int epi = EPI_STORE_BF16;
  const float* bias = nullptr;
  const __nv_bfloat16* residual = nullptr; int ldr = 0;
  const float* pos = nullptr; int pos_period = 0;
  const __nv_bfloat16* pos_tile = nullptr;   // optional [128, N] bf16 copy of `pos` tiled to 128 rows (pos_period must divide 128):
                                             // lets the 2-CTA GEMM add it through its shared-memory-staged epilogue
  float out_scale = 1.0f;
  bool allow_2cta = true;     // per-call switch (VitmarlVitOptions::gemm_2cta / vitmarl_gemm_bf16 flags): false = 1-CTA kernels only
  float* colsum_a = nullptr;   // MN-major A only (dW = A^T . B): also accumulate the column sums of A (the bias gradient), fp32 [M];
                               // fused into the weight-gradient kernel when it takes the product, a separate pass otherwise
};

int num_sms();
int launch_gemm(cudaStream_t stream, const GemmDesc& g);
// 2-CTA (cta_group::2, M = 256 per CTA pair) variant for K-major operands; used by launch_gemm when supported
bool gemm2_supported(const GemmDesc& g);
int launch_gemm2(cudaStream_t stream, const GemmDesc& g);
// dW-shaped products (MN x MN, split-K, fp32 red.add) on CTA pairs with 256 x 384 work items; returns 1 when the shape is not handled
int launch_gemm2_dw(cudaStream_t stream, const GemmDesc& g);

// x[B,H,W,C] bf16 -> patches [B*T, P*P*C] bf16, token t = py*(W/P)+px, feature (ph, pw, c)
int launch_patchify(cudaStream_t s, const __nv_bfloat16* x, __nv_bfloat16* out, int B, int H, int W, int C, int P);
// d(patches) -> d(x) (exact inverse scatter; every pixel belongs to one patch)
int launch_unpatchify(cudaStream_t s, const __nv_bfloat16* dpatches, __nv_bfloat16* dx, int B, int H, int W, int C, int P);

// y = LN(x) * gamma + beta over the last dim D (bf16 in/out, fp32 statistics); stats[M,2] = (mean, rstd) or null
int launch_layernorm(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y,
                     float* stats, int M, int D, float eps);
// LN backward: dx (+)= dLN(dy); dgamma/dbeta accumulated with atomics (fp32, pre-zeroed)
int launch_layernorm_bwd(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* stats, const __nv_bfloat16* dy,
                         const __nv_bfloat16* dx_add, __nv_bfloat16* dx, float* dgamma, float* dbeta, int M, int D);

// multi-head self-attention over T = 64 tokens, head dim 64: qkv [B*64, 3*D] (q | k | v, head-major) -> out [B*64, D]
int launch_attention(cudaStream_t s, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int heads);
// dbqkv [3*D] fp32 (nullable): += column sums of dqkv (the QKV bias gradient), reduced inside the kernel
int launch_attention_bwd(cudaStream_t s, const __nv_bfloat16* qkv, const __nv_bfloat16* dout, __nv_bfloat16* dqkv, int B, int heads,
                         float* dbqkv = nullptr);

// y[B,D] (fp32) = mean_t LN(x[B*T,D]); saves stats[M,2]
int launch_final_ln_pool(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* beta, float* y, float* stats,
                         int B, int T, int D, float eps);
// dx[B*T,D] (bf16) from dy[B,D]; dgamma/dbeta atomics
int launch_final_ln_pool_bwd(cudaStream_t s, const __nv_bfloat16* x, const float* gamma, const float* stats, const float* dy,
                             __nv_bfloat16* dx, float* dgamma, float* dbeta, int B, int T, int D);

// Per-launch switches of the fused block kernels (no process-global state: the C ABI passes them per call)
struct FusedOpts {
  bool pdl = true;            // programmatic dependent launch
  int attn_flags = 4;         // FusedAttn2Params::flags
  long long* dbg = nullptr;   // device buffer (>= 512 int64) for clock64 phase stamps: [0,256) MLP block, [256,512) attention block
};

// CTA-pair (cta_group::2) fused MLP block on FOLDED parameters (vit_fold.cu): w1f = W1.diag(gamma), b1p = bf16(b1 + W1.beta),
// w2h = W2/2.  out = x + fc2(gelu(fc1(LN(x))));  `out` may alias `x`
int launch_fused_mlp2(cudaStream_t stream, const __nv_bfloat16* x, __nv_bfloat16* out, const __nv_bfloat16* w1f, const __nv_bfloat16* b1p,
                      const __nv_bfloat16* w2h, const float* b2, int M, int D, int hidden, float eps, const FusedOpts& o);
bool fused_mlp2_supported(int D, int hidden);

// Fused BACKWARD of the MLP block on the same folded parameters (fused_mlp_bwd.cu): from the block input x and dy = d(out) it
// recomputes the block and writes xhat [M,D], h2 = 2 gelu(hpre) [M,hidden], dh = d(hpre) [M,hidden] (operands of the two
// weight-gradient products) and dx [M,D] (may alias dy).
int launch_fused_mlp_bwd(cudaStream_t stream, const __nv_bfloat16* x, const __nv_bfloat16* dy, const __nv_bfloat16* w1f, const __nv_bfloat16* b1p,
                         const __nv_bfloat16* w2h, __nv_bfloat16* xhat, __nv_bfloat16* h2, __nv_bfloat16* dh, __nv_bfloat16* dx, int M, int D,
                         int hidden, float eps, float* dbf, float* dbx, long long* dbg = nullptr);
bool fused_mlp_bwd_supported(int D, int hidden);
// gradients of folded parameters W' = c W diag(gamma), b' = c (b + W beta) back to (W, b, gamma, beta), all accumulated (+=):
// dWf [N,K] / dbf [N] fp32 are the gradients w.r.t. W' / b' (dbf, db, dbeta may be null; gamma null = ones)
int launch_unfold_grads(cudaStream_t s, int N, int K, float c, const __nv_bfloat16* W, const float* gamma, const float* beta, const float* dWf,
                        const float* dbf, float* dW, float* db, float* dgamma, float* dbeta);

// Second-generation fused attention block on FOLDED parameters (vit_fold.cu): wqkvf = Wqkv.diag(gamma) with the Q rows scaled
// by log2(e)/8, bqp = bf16 folded Q bias, bof = bo + Wo.(bv + Wv.beta).  `out` may alias `x`
int launch_fused_attn2(cudaStream_t stream, const __nv_bfloat16* x, __nv_bfloat16* out, const __nv_bfloat16* wqkvf, const __nv_bfloat16* bqp,
                       const __nv_bfloat16* wo, const float* bof, int M, int D, int heads, float eps, const FusedOpts& o);
bool fused_attn2_supported(int D, int heads, int tokens);

// parameter folding: Wf = scale * W . diag(gamma) (bf16 [N,K]), bias_out = bias + W . beta (fp32 or bf16 [N]); null = identity
struct FoldJob {
  const __nv_bfloat16* W; const float* bias; const float* gamma; const float* beta;
  __nv_bfloat16* Wf; void* bias_out;
  int N, K, bias_out_bf16;
  float scale;
};
struct FoldJobs { FoldJob job[64]; int n; };
int launch_fold_params(cudaStream_t s, const FoldJobs& jobs);
int launch_pos_tile(cudaStream_t s, const float* pos, __nv_bfloat16* out, int T, int D);   // [T,D] fp32 -> [128,D] bf16 (T | 128)

// elementwise helpers
int launch_gelu_bwd(cudaStream_t s, const __nv_bfloat16* pre, const __nv_bfloat16* dy, __nv_bfloat16* dx, size_t n);   // dx = dy * gelu'(pre)
int launch_colsum(cudaStream_t s, const __nv_bfloat16* x, float* out, int M, int N);                                  // out[N] += sum_m x[m,n]
int launch_cast_f32_to_bf16(cudaStream_t s, const float* src, __nv_bfloat16* dst, size_t n);
int launch_transpose_f32_to_bf16(cudaStream_t s, const float* src, __nv_bfloat16* dst, int rows, int cols);            // dst[cols, rows] = src[rows, cols]^T

}  // namespace vitmarl
