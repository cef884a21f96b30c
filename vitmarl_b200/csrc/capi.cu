// C-ABI plumbing shared by all entry points: version, thread-local error text.
#include <cuda_runtime.h>
#include <stdio.h>

#include "common.cuh"
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
