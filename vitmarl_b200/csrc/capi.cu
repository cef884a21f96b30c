// C-ABI plumbing shared by all entry points: version, thread-local error text.
#include <cuda_runtime.h>
#include <stdio.h>

#include "common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

static thread_local char g_err[512] = "";

const char* set_last_error(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
  return g_err;
}

int check_cuda(cudaError_t err) {
  if (err == cudaSuccess) return VITMARL_OK;
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s): %s", (int)err, cudaGetErrorName(err), cudaGetErrorString(err));
  return VITMARL_ECUDA;
}

}  // namespace vitmarl

extern "C" int vitmarl_abi_version(void) { return VITMARL_ABI_VERSION; }
extern "C" const char* vitmarl_last_error(void) { return vitmarl::g_err; }

extern "C" int vitmarl_gemm_bf16(void* stream, int M, int N, int K, const void* A, int lda, int a_mn_major, const void* B,
                                 int ldb, int b_mn_major, void* C, int ldc, int epi, const float* bias, const void* residual,
                                 int ldr, const float* pos, int pos_period, float out_scale, int flags) {
  vitmarl::GemmDesc g;
  g.M = M; g.N = N; g.K = K;
  g.A = static_cast<const __nv_bfloat16*>(A); g.lda = lda; g.a_mn_major = a_mn_major != 0;
  g.B = static_cast<const __nv_bfloat16*>(B); g.ldb = ldb; g.b_mn_major = b_mn_major != 0;
  g.C = C; g.ldc = ldc; g.epi = epi; g.bias = bias;
  g.residual = static_cast<const __nv_bfloat16*>(residual); g.ldr = ldr; g.pos = pos; g.pos_period = pos_period;
  g.out_scale = out_scale; g.allow_2cta = !(flags & VITMARL_GEMM_NO_2CTA);
  if (epi < 0 || epi > 4) return VITMARL_EINVAL;
  return vitmarl::launch_gemm(static_cast<cudaStream_t>(stream), g);
}

extern "C" int vitmarl_attention_fwd(void* stream, int B, int heads, const void* qkv, void* out) {
  if (!qkv || !out) return VITMARL_EINVAL;
  return vitmarl::launch_attention(static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), B, heads);
}

extern "C" int vitmarl_attention_bwd(void* stream, int B, int heads, const void* qkv, const void* dout, void* dqkv) {
  if (!qkv || !dout || !dqkv) return VITMARL_EINVAL;
  return vitmarl::launch_attention_bwd(static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout),
                                       static_cast<__nv_bfloat16*>(dqkv), B, heads);
}
