// Microbenchmark: sustained issue/execute cost of tcgen05.mma (kind::f16, bf16, M = 128 per CTA) as a function of
// N, operand source (A in shared memory = SS, A in tensor memory = TS) and cta_group (1 or 2).  One CTA (pair) per
// SM (pair), one thread issues `iters` x 4 MMAs back to back (accumulating into the same TMEM tile), then commits and
// waits.  Operand contents are irrelevant (uninitialised smem / TMEM).  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu -I ../../vitmarl_b200/csrc
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

namespace vitmarl {
const char* set_last_error(const char* m) { return m; }
int check_cuda(cudaError_t e) { return e == cudaSuccess ? 0 : -3; }

template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if constexpr (CG == 2) umma_bf16_2sm(d, a, b, idesc, 1u); else umma_bf16(d, a, b, idesc, 1u);
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  if constexpr (CG == 2) umma_bf16_2sm_ts(d, a, b, idesc, 1u);
  else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
}

template <int CG, int N, bool TS, int BSTAGES>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t barA = sbase + 200 * 1024, slot = barA + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (CG == 2) rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(barA, 1); fence_mbar_init(); }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc<512>(slot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 0 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(CG * 128, N, false, false);
    const uint32_t la = umma_desc_lo(sbase);                    // A: [128 x 64] K-block (16 KB) at 0
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t lb = umma_desc_lo(sbase + 16384 + (i % BSTAGES) * 32768);   // B: [<=256 rows x 64] K-block, rotating stages
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (TS) mma_ts<CG>(tmem_base + 256, tmem_base + 8 * k + 32 * (i & 3), umma_desc_from_lo(lb + 2 * k), idesc);
          else mma_ss<CG>(tmem_base + 256, umma_desc_from_lo(la + 2 * k), umma_desc_from_lo(lb + 2 * k), idesc);
        }
      }
      const long long ti = clock64();
      if constexpr (CG == 2) umma_commit_2sm(barA); else umma_commit(barA);
      mbar_wait(barA, 0);
      t1 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    }
    __syncwarp();
  } else if (warp == 0) {
    mbar_wait(barA, 0);            // follower: the multicast commit arrives here too
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else tmem_dealloc<512>(tmem_base);
  }
}

template <int CG, int N, bool TS, int BSTAGES>
void run(const char* name, long long* d_out) {
  const int iters = 512, smem = 206 * 1024;
  auto k = mma_rate_kernel<CG, N, TS, BSTAGES>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, iters, d_out);
    if (e != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: failed %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h[0] / (iters * 4), per_issue = (double)h[1] / (iters * 4);
  const double floor_c = 128.0 * N / 256.0;
  printf("%-28s N=%3d  %7.1f cyc/MMA (issue %6.1f)  floor %5.1f  eff %.2f\n", name, N, per, per_issue, floor_c, floor_c / per);
}
}  // namespace vitmarl

int main() {
  using namespace vitmarl;
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<1, 64, false, 1>("cg1 SS", d_out);   run<1, 128, false, 1>("cg1 SS", d_out);  run<1, 192, false, 1>("cg1 SS", d_out);  run<1, 256, false, 1>("cg1 SS", d_out);
  run<1, 64, true, 1>("cg1 TS", d_out);    run<1, 128, true, 1>("cg1 TS", d_out);   run<1, 192, true, 1>("cg1 TS", d_out);   run<1, 256, true, 1>("cg1 TS", d_out);
  run<2, 64, false, 1>("cg2 SS", d_out);   run<2, 128, false, 1>("cg2 SS", d_out);  run<2, 192, false, 1>("cg2 SS", d_out);  run<2, 256, false, 1>("cg2 SS", d_out);
  run<2, 64, true, 1>("cg2 TS", d_out);    run<2, 128, true, 1>("cg2 TS", d_out);   run<2, 192, true, 1>("cg2 TS", d_out);   run<2, 256, true, 1>("cg2 TS", d_out);
  run<2, 128, false, 4>("cg2 SS 4 B stages", d_out); run<2, 192, true, 4>("cg2 TS 4 B stages", d_out);
  return 0;
}
