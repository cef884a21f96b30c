// Microbenchmark: does tcgen05.ld (epilogue traffic) slow down under a concurrent tcgen05.mma stream, and vice versa?
// One CTA per SM: warp 0 issues `iters` x 4 MMAs (cta_group::1, M=128, N, SS or TS); warps 4..19 loop
// { tcgen05.ld 32x32b.x32 ; wait ; 8 dependent FADDs } on their lane quadrant until the MMA stream has retired.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

namespace vitmarl {
const char* set_last_error(const char* m) { return m; }
int check_cuda(cudaError_t e) { return e == cudaSuccess ? 0 : -3; }

template <int N, bool TS, bool WITH_MMA>
__global__ void __launch_bounds__(640, 1) kernel(int iters, int ld_warps, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t barA = sbase + 100 * 1024, slot = barA + 8;
  volatile int* done = reinterpret_cast<volatile int*>(smem_raw + (sbase - smem_u32(smem_raw)) + 100 * 1024 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(barA, 1); fence_mbar_init(); *done = 0; }
  if (warp == 1) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
    const uint32_t la = umma_desc_lo(sbase), lb = umma_desc_lo(sbase + 16384);
    const long long t0 = clock64();
    if (WITH_MMA) {
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if constexpr (TS) umma_bf16_ts(tmem_base + 256, tmem_base + 8 * k + 32 * (i & 3), umma_desc_from_lo(lb + 2 * k), idesc, 1u);
            else umma_bf16(tmem_base + 256, umma_desc_from_lo(la + 2 * k), umma_desc_from_lo(lb + 2 * k), idesc, 1u);
          }
        }
        umma_commit(barA);
      }
      __syncwarp();
      mbar_wait(barA, 0);
    } else {
      while (clock64() - t0 < 200000) {}
    }
    const long long t1 = clock64();
    *done = 1;
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  } else if (warp >= 4 && warp < 4 + ld_warps) {
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 128 + ((warp - 4) >> 2) * 32 % 128;   // columns 128..255 (not the MMA's D / A)
    uint32_t r[32];
    float acc = 0.f;
    long long n = 0;
    const long long t0 = clock64();
    while (!*done) {
      tmem_ld_32x32(taddr, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += __uint_as_float(r[4 * i]);
      ++n;
    }
    const long long t1 = clock64();
    if (acc == 123.456f) sink[0] = acc;
    if (blockIdx.x == 0 && lane == 0) { out[1 + (warp - 4) * 2] = n; out[2 + (warp - 4) * 2] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

template <int N, bool TS, bool WITH_MMA>
void run(const char* name, int ld_warps, long long* d_out, float* sink) {
  const int iters = 1024, smem = 104 * 1024;
  auto k = kernel<N, TS, WITH_MMA>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d_out, 0, 64 * 8);
  k<<<148, 640, smem>>>(iters, ld_warps, d_out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: failed %s\n", name, cudaGetErrorString(e)); return; }
  long long h[64];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double lds = 0, cyc = 1;
  for (int w = 0; w < ld_warps; ++w) { lds += h[1 + 2 * w]; cyc = h[2 + 2 * w]; }
  printf("%-10s N=%3d ld_warps=%2d : %6.1f cyc/MMA | ld: %7.1f cyc per ld per warp, %7.1f B/clk/SM\n", name, N,
         ld_warps, WITH_MMA ? (double)h[0] / (iters * 4) : 0.0, ld_warps ? cyc * ld_warps / lds : 0.0, lds * 4096.0 / cyc);
}
}  // namespace vitmarl

int main() {
  using namespace vitmarl;
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 64 * 8); cudaMalloc(&sink, 64);
  run<128, false, false>("no MMA", 4, d_out, sink);  run<128, false, false>("no MMA", 16, d_out, sink);
  run<128, false, true>("SS", 0, d_out, sink); run<128, false, true>("SS", 4, d_out, sink); run<128, false, true>("SS", 16, d_out, sink);
  run<192, false, true>("SS", 16, d_out, sink);
  run<128, true, true>("TS", 0, d_out, sink);  run<128, true, true>("TS", 4, d_out, sink);  run<128, true, true>("TS", 16, d_out, sink);
  run<192, true, true>("TS", 16, d_out, sink); run<64, true, true>("TS", 16, d_out, sink);
  return 0;
}
