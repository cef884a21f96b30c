// Microbenchmark: tcgen05.ld / tcgen05.st throughput (32x32b.x32: 4 KB per warp instruction) with W warps per CTA,
// one CTA per SM.  Prints cycles per warp-instruction and bytes/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_rate tmem_rate.cu -I ../../vitmarl_b200/csrc
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

namespace vitmarl {
const char* set_last_error(const char* m) { return m; }
int check_cuda(cudaError_t e) { return e == cudaSuccess ? 0 : -3; }

template <int MODE>   // 0: ld x32, 1: st x32, 2: ld x32 + 32 FADDs consuming the data
__global__ void __launch_bounds__(1024, 1) tmem_rate_kernel(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  __shared__ long long tmin, tmax;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { tmin = 0x7fffffffffffffffLL; tmax = 0; }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) * 32) % 512;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = lane + i;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1) {
      tmem_st_32x32(taddr, r);
    } else {
      tmem_ld_32x32(taddr, r);
      if (MODE == 2) {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
      }
    }
  }
  if (MODE == 1) tmem_st_wait(); else tmem_ld_wait();
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
  if (acc == 123.456f) sink[0] = acc;
  if (lane == 0) { atomicMin(&tmin, t0); atomicMax(&tmax, t1); }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = tmax - tmin;
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

template <int MODE>
void run(const char* name, int warps, long long* d_out, float* sink) {
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) {
    tmem_rate_kernel<MODE><<<148, warps * 32>>>(iters, d_out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: failed %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h;
  cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h / iters;
  printf("%-14s warps=%2d  %7.1f cyc per round (all warps 1 instr each)  %7.1f B/clk/SM  %6.1f cyc/instr/SMSP\n", name, warps, per,
         warps * 4096.0 / per, per / ((warps + 3) / 4));
}
}  // namespace vitmarl

int main() {
  using namespace vitmarl;
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&sink, 64);
  for (int w : {4, 8, 16, 32}) run<0>("ld.x32", w, d_out, sink);
  for (int w : {4, 8, 16, 32}) run<1>("st.x32", w, d_out, sink);
  for (int w : {4, 8, 16, 32}) run<2>("ld.x32+32FADD", w, d_out, sink);
  return 0;
}
