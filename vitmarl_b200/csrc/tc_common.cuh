// Blackwell tensor-core plumbing for sm_100a written as inline PTX: TMA tiled loads,
// UMMA (tcgen05.mma) shared-memory / instruction descriptors, TMEM alloc + loads, commit.
// Bit layouts follow the PTX ISA "tcgen05" matrix/instruction descriptor tables (cross-checked
// against CuTe's cute/arch/mma_sm100_desc.hpp, which is used as documentation only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "common.cuh"

namespace vitmarl {

// ---- host: tensor maps (driver entry point fetched at run time; no libcuda link dependency) ----
// 2-D bf16 row-major tensor [rows, cols] (cols contiguous); box = [box_rows, box_cols]; 128B swizzle
// requires box_cols * 2 bytes == 128.
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                      uint32_t box_rows, uint32_t box_cols);

#ifdef __CUDACC__
// ---- TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tile load: c0 = innermost (column) coordinate, c1 = row coordinate
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(m), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tile store smem -> global (bulk async group); rows beyond the tensor extent are clipped by the TMA unit
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}

// 2-D tile reduction smem -> global: global[tile] += smem[tile], element type from the tensor map (bf16: the add and its
// rounding happen in L2).  `x += delta` of a residual stream updated in place needs no residual load at all.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(src_smem), "r"(c0),
               "r"(c1)
               : "memory");
}

// ---- UMMA descriptors -------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, tile rows are 128 bytes (64 bf16), 8-row atoms of 1024 B.
//   K-major  (rows = M/N index, 128 B of K per row):  LBO unused (=1), SBO = 1024 B between 8-row groups.
//   MN-major (rows = K index, 128 B = 64 MN elements per row): SBO = 1024 B between 8-K groups,
//            LBO = bytes between consecutive 64-element MN blocks.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor: kind::f16, A/B = bf16, D = f32, M x N, optional MN-major operands.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The descriptor's upper word is constant for SWIZZLE_128B / SBO = 1024; the lower word is (addr >> 4) | LBO << 16, so a
// K = 16 step inside a 128-byte row (+32 bytes) is `lo + 2`: issue loops build `lo` once per K-block and add immediates
// (the generic constructor above costs ~5 dependent uniform-datapath ops per descriptor on the single issuing thread).
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes = 16) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t umma_desc_from_lo(uint32_t lo, uint32_t sbo_bytes = 1024) {
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// One lane of a converged warp (warp-uniform predicate: the compiler emits a plain predicated instruction stream instead
// of the per-active-thread loop it generates under `if (lane == 0)`).
// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while the previous kernel of
// the stream is still running; griddep_wait() blocks the calling thread until that kernel has completed and its
// writes are visible (it returns at once for a normal launch).  Discipline used here: every thread that touches
// memory the previous kernel reads or writes waits first, and a kernel lets ITS dependent go only after its own
// wait, so a kernel never overlaps anything older than its immediate predecessor.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (TS form): A[m, k] at lane m, 32-bit column a_tmem + k/2 (bf16 pairs)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {   // whole warp; COLS power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {    // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives lane (base_lane + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns register -> TMEM store (thread i of the warp writes lane base_lane + i, columns [c, c+16))
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// the same 32-bit word to 16 consecutive columns of the thread's lane (zero fill)
__device__ __forceinline__ void tmem_st_32x16_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2) ------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst_smem, const CUtensorMap* m, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(m), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0)
      : "memory");
}
// A operand in tensor memory (TS form): A[m, k] of the CTA's 128 rows at lane m, 32-bit column a_tmem + k/2 (bf16 pairs)
__device__ __forceinline__ void umma_bf16_2sm_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive (default .release.cta, as CUTLASS' ClusterBarrier::arrive(cta_id)) on an mbarrier given by its shared::cluster address (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// mbarrier wait with a watchdog: a protocol bug traps after ~2 s instead of hanging the GPU
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires) instead
// of spinning through the ~100-cycle default time-out -- the polling loops of 20+ waiting warps were ~20-30 % of the
// fused kernels' issued instructions, taken from the issue slots of the warps that work
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Build-time experiment knobs (the sustained rollout loop runs at the GPU's power cap, so polling costs clock):
//   VITMARL_WAIT_SLEEP_NS  > 0 : __nanosleep between failed polls;  VITMARL_WAIT_SLEEP_ALL = 0 : only in warps >= 4 (the epilogue /
//   softmax / conversion roles of the fused kernels; warps 0-3 are the TMA and MMA-issuing roles whose wake-up latency is critical)
#ifndef VITMARL_WAIT_SLEEP_NS
#define VITMARL_WAIT_SLEEP_NS 0
#endif
#ifndef VITMARL_WAIT_SLEEP_ALL
#define VITMARL_WAIT_SLEEP_ALL 0
#endif
#ifndef VITMARL_WAIT_HINT
#define VITMARL_WAIT_HINT 20000u      // suspend-time hint of mbarrier.try_wait
#endif
__device__ __forceinline__ void mbar_wait_guard(uint32_t bar, uint32_t parity) {
  uint32_t polls = 0;
  long long t0 = 0;
  while (!mbar_try_wait_hint(bar, parity, VITMARL_WAIT_HINT)) {
    if (VITMARL_WAIT_SLEEP_NS > 0 && (VITMARL_WAIT_SLEEP_ALL || threadIdx.x >= 128)) __nanosleep(VITMARL_WAIT_SLEEP_NS);
    if ((++polls & 63u) == 0) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) {
#ifdef VITMARL_WATCHDOG_PRINT   // names the barrier that hung (debug builds only: the printf costs a stack frame and spills)
        if ((threadIdx.x & 31) == 0) printf("vitmarl watchdog: block %d warp %d stuck on mbarrier +0x%x parity %u\n", blockIdx.x, threadIdx.x >> 5, bar & 0x3ff, parity);
#endif
        __trap();
      }
    }
  }
}

// (Measured: letting ONE lane poll and parking the other 31 at __syncwarp() is ~1.7x SLOWER end to end than all 32 lanes
// executing try_wait -- the divergent spin delays the warp's wake-up -- so waits are always executed by the whole warp.)

// ---- small math -------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// flax nn.gelu default (approximate=True): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanh_approx(u));
}
// 8 FP32 instructions + 1 MUFU (the epilogue that applies it is instruction-bound: 21 warp-instructions per element measured with
// the textbook form): g' = 0.5 + 0.5 (t + x (1 - t^2) k0 (1 + 3 k1 x^2)),  t = tanh(x k0 (1 + k1 x^2))
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float s = x * x;
  const float t = tanh_approx(x * fmaf(s, k0k1, k0));
  const float c = fmaf(x * fmaf(-t, t, 1.0f), fmaf(s, 3.0f * k0k1, k0), t);
  return fmaf(c, 0.5f, 0.5f);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// gelu_tanh on a packed bf16x2 pair: 5 packed FP ops + 1 packed MUFU for two elements
// (x^2, fma, mul, tanh.approx.bf16x2, mul, fma) -- the epilogue of the fused MLP is issue-bound on the CUDA cores.
__device__ __forceinline__ uint32_t gelu_tanh_bf16x2(uint32_t x) {
  const uint32_t C0 = 0x3F4C3F4Cu;   // bf16(0.7978845608) x2
  const uint32_t C1 = 0x3D123D12u;   // bf16(0.7978845608 * 0.044715 = 0.0356774) x2
  const uint32_t HALF = 0x3F003F00u; // bf16(0.5) x2
  uint32_t x2, p, u, t, hx, y;
  asm("mul.rn.bf16x2 %0, %1, %1;" : "=r"(x2) : "r"(x));
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(x2), "r"(C1), "r"(C0));
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(u) : "r"(p), "r"(x));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(u));
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(hx) : "r"(x), "r"(HALF));
  asm("fma.rn.bf16x2 %0, %1, %2, %1;" : "=r"(y) : "r"(hx), "r"(t));
  return y;
}
// ---- packed bf16x2 / mixed-precision helpers (HFMA2.BF16, FHADD.BF16, FHFMA.BF16) ----
__device__ __forceinline__ uint32_t bf16x2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf16x2_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// mixed-precision accumulate: s += lo + hi, q += lo^2 + hi^2 for a packed bf16 pair (FHADD / FHFMA, no unpacking)
__device__ __forceinline__ void stats_bf16x2(uint32_t w, float& s, float& q) {
  asm("{\n\t.reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %2;\n\t"
      "add.f32.bf16 %0, lo, %0;\n\t"
      "add.f32.bf16 %0, hi, %0;\n\t"
      "fma.rn.f32.bf16 %1, lo, lo, %1;\n\t"
      "fma.rn.f32.bf16 %1, hi, hi, %1;\n\t}"
      : "+f"(s), "+f"(q)
      : "r"(w));
}
// (x - mean) * rstd for a packed pair: fp32 FMAs with bf16 multiplicands (x, rstd_b), fp32 addend -mean * rstd_b
__device__ __forceinline__ uint32_t norm_bf16x2(uint32_t w, uint32_t rstd_b, float nmr) {
  float a, b;
  asm("{\n\t.reg .b16 lo, hi, r;\n\t"
      "mov.b32 {lo, hi}, %2;\n\t"
      "mov.b32 {r, _}, %3;\n\t"
      "fma.rn.f32.bf16 %0, lo, r, %4;\n\t"
      "fma.rn.f32.bf16 %1, hi, r, %4;\n\t}"
      : "=f"(a), "=f"(b)
      : "r"(w), "r"(rstd_b), "f"(nmr));
  return pack_bf16(a, b);
}

// Column sums over the 32 lanes of a warp for 32 per-lane values (lane = row, v[i] = column i): a transposing butterfly --
// 31 shuffles + 31 adds instead of 32 x 5; afterwards lane l holds sum_rows v[l] in v[0].
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// the same for 16 per-lane values: lanes l and l + 16 both end with the sum over all 32 rows of column l & 15
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
#endif  // __CUDACC__

}  // namespace vitmarl
