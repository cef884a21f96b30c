// Microbenchmark: issue rate of the CUDA-core instructions the fused epilogues are made of (per SM sub-partition):
// FFMA, HFMA2.BF16 (packed), F2FP (cvt.rn.bf16x2.f32), FHFMA.BF16 (mixed precision), MUFU.TANH.BF16, MUFU.EX2.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* out, float* sink) {
  float f[8]; uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 0.001f + i; u[i] = 0x3f803f80u + threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f[i]) : "f"(1.0001f));
      if (OP == 1) asm volatile("fma.rn.bf16x2 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(0x3f803f80u));
      if (OP == 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(f[(i + 1) & 7]));
      if (OP == 3) asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.bf16 %0, lo, hi, %0;\n\t}" : "+f"(f[i]) : "r"(u[i]));
      if (OP == 4) asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %0;\n\ttanh.approx.bf16 lo, lo;\n\tmov.b32 %0, {lo, hi};\n\t}" : "+r"(u[i]));
      if (OP == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 6) asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(0x3c003c00u));
      if (OP == 7) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(u[i]);
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP> void run(const char* name, long long* d, float* sink) {
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    k<OP><<<148, warps * 32>>>(iters, d, sink); cudaDeviceSynchronize();
    k<OP><<<148, warps * 32>>>(iters, d, sink); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = (double)h / (iters * 8.0 * (warps / 4));   // cycles per warp-instruction per sub-partition
    printf("%-22s warps/SM=%2d : %6.2f cycles per warp-instr per SMSP  (%5.1f lanes/clk/SM)\n", name, warps, per_smsp, 128.0 / per_smsp);
  }
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 64); cudaMalloc(&sink, 64);
  run<0>("FFMA", d, sink); run<1>("HFMA2.BF16 (packed)", d, sink); run<6>("HADD2.BF16 (packed)", d, sink); run<2>("F2FP bf16x2<-f32", d, sink);
  run<3>("FHFMA.BF16 (mixed)", d, sink); run<4>("MUFU.TANH.BF16", d, sink); run<7>("MUFU.TANH.F32", d, sink); run<5>("MUFU.EX2", d, sink);
  return 0;
}
