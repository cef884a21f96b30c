// Fused transformer MLP block for D = 192 (ViT-Tiny), inference path:
//     out = x + FC2( gelu_tanh( FC1( LayerNorm(x) ) ) )
// ONE kernel per layer; the [tokens, 4D] hidden activation never leaves the SM (the unfused v0 wrote
// and re-read 2 x 403 MB per layer at 4096 images, plus a separate LayerNorm pass).
//
// Per CTA: a 128-token tile (persistent over tiles).  Warp roles:
//   warp 0      TMA producer: x tile (3 K-blocks, 128B swizzle) + weight K-blocks in MMA order
//   warp 1      TMEM allocator + tcgen05.mma issuer
//   warps 2-9   compute: LayerNorm in place in shared memory (2 threads per token row), GELU epilogue
//               TMEM -> registers -> bf16 -> swizzled shared memory (the A operand of FC2), final
//               bias + residual epilogue to global memory
// Hidden dimension processed in 6 chunks of 128:  acc1[c&1] = LN(x) . W1[c]^T  (TMEM, double buffered),
// H[c&1] = gelu(acc1 + b1) (smem, double buffered), acc2 += H[c&1] . W2[:,c]^T  (TMEM, 192 columns).
// Software-pipelined issue order: FC1(0) FC1(1) FC2(0) FC1(2) FC2(1) ... FC2(5), so the tensor pipe works on
// FC1(c+1) while the compute warps run GELU on chunk c.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

namespace fmlp {
constexpr int D = 192, HID = 768, HC = 128, NCHUNK = HID / HC;   // 6 hidden chunks of 128 (a tcgen05.mma costs ~100+ cycles
                                                                  // whatever N <= 128 is, so wide chunks halve the FC1 issue time)
constexpr int TM = 128;                                           // tokens per tile
constexpr int KB_X = D / 64;                                      // 3 K-blocks of x / LN(x)
constexpr int KB_H = HC / 64;                                      // 2 K-blocks per hidden chunk
constexpr int NB = 2;                                             // acc1 / H buffers
constexpr int NS1 = 3;                                            // W1 ring: K-blocks [128 x 64] (16 KB) = 1 chunk
constexpr int NS2 = 2;                                            // W2 ring: K-blocks [192 x 64] (24 KB) = 1 chunk
constexpr int S1_BYTES = HC * 128, S2_BYTES = D * 128;
constexpr int XN_BYTES = KB_X * TM * 128;                         // 48 KB
constexpr int H_BYTES = KB_H * TM * 128;                          // 32 KB per buffer
constexpr int OFF_XN = 0;
constexpr int OFF_H = OFF_XN + XN_BYTES;
constexpr int OFF_W1 = OFF_H + NB * H_BYTES;
constexpr int OFF_W2 = OFF_W1 + NS1 * S1_BYTES;
constexpr int OFF_BAR = OFF_W2 + NS2 * S2_BYTES;
constexpr int OFF_MISC = OFF_BAR + 512;                           // LN partials [128][2] float2, gamma/beta, biases
constexpr int MISC_BYTES = 128 * 4 * 8 + (2 * D + HID + D) * 4;
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int NCW = 16;                                           // compute warps: 4 per TMEM lane quadrant, each owns a quarter of the columns
constexpr int THREADS = 96 + 32 * NCW;                            // 3 control warps + 16 compute warps
constexpr int TMEM_COLS = 512;                                    // acc1[2] @ 0,128 ; acc2 @ 256 (192 cols)
constexpr int ACC2_COL = 256;
// barrier slots (8 bytes each)
enum { B_XFULL = 0, B_XEMPTY, B_XNREADY, B_ACC2FULL, B_ACC2EMPTY, B_ACC1FULL = 5, B_ACC1EMPTY = B_ACC1FULL + NB,
       B_HREADY = B_ACC1EMPTY + NB, B_HEMPTY = B_HREADY + NB, B_W1FULL = B_HEMPTY + NB, B_W1EMPTY = B_W1FULL + NS1,
       B_W2FULL = B_W1EMPTY + NS1, B_W2EMPTY = B_W2FULL + NS2, B_RESFULL = B_W2EMPTY + NS2, B_RESREAD, B_TMEMSLOT, B_COUNT };
static_assert(B_COUNT * 8 <= 512, "barrier area");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace fmlp

struct FusedMlpParams {
  int M;                        // tokens
  const __nv_bfloat16* x;       // [M, D] residual stream in
  __nv_bfloat16* out;           // [M, D] (may alias x)
  const float* gamma; const float* beta;   // LayerNorm
  const float* b1; const float* b2;
  float eps;
  long long* dbg;               // optional timeline buffer (clock64 stamps of CTA 0, second tile); null in production
};

#define FMLP_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && it == 1) p.dbg[(slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(fmlp::THREADS, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const FusedMlpParams p) {
  using namespace fmlp;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  auto bar = [&](int i) { return sbase + OFF_BAR + 8u * i; };
  float2* ln_part = reinterpret_cast<float2*>(sptr + OFF_MISC);                    // [128][4]
  float* s_gamma = reinterpret_cast<float*>(sptr + OFF_MISC + 128 * 4 * 8);
  float* s_beta = s_gamma + D;
  float* s_b1 = s_beta + D;
  float* s_b2 = s_b1 + HID;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M + TM - 1) / TM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmOut);
    mbar_init(bar(B_XFULL), 1); mbar_init(bar(B_XEMPTY), 1); mbar_init(bar(B_XNREADY), NCW);
    mbar_init(bar(B_ACC2FULL), 1); mbar_init(bar(B_ACC2EMPTY), NCW); mbar_init(bar(B_RESFULL), 1); mbar_init(bar(B_RESREAD), NCW);
    for (int i = 0; i < NB; ++i) {
      mbar_init(bar(B_ACC1FULL + i), 1); mbar_init(bar(B_ACC1EMPTY + i), NCW);
      mbar_init(bar(B_HREADY + i), NCW); mbar_init(bar(B_HEMPTY + i), 1);
    }
    for (int i = 0; i < NS1; ++i) { mbar_init(bar(B_W1FULL + i), 1); mbar_init(bar(B_W1EMPTY + i), 1); }
    for (int i = 0; i < NS2; ++i) { mbar_init(bar(B_W2FULL + i), 1); mbar_init(bar(B_W2EMPTY + i), 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(bar(B_TMEMSLOT));
  for (int i = threadIdx.x; i < D; i += THREADS) { s_gamma[i] = p.gamma[i]; s_beta[i] = p.beta[i]; s_b2[i] = p.b2[i]; }
  for (int i = threadIdx.x; i < HID; i += THREADS) s_b1[i] = p.b1[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bar(B_TMEMSLOT)));

  if (warp == 0) {
    // =============================== TMA producer: x tiles + W1 K-blocks ===============================
    if (lane == 0) {
      int s1 = 0; uint32_t ph1 = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(bar(B_RESREAD), (it & 1) ^ 1);            // previous tile's epilogue has read its residual out of the x buffer
        mbar_arrive_expect_tx(bar(B_XFULL), XN_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_XN + kb * TM * 128, &tmX, kb * 64, tile * TM, bar(B_XFULL));
        for (int c = 0; c < NCHUNK; ++c)
          for (int kb = 0; kb < KB_X; ++kb) {
            mbar_wait(bar(B_W1EMPTY + s1), ph1 ^ 1);
            mbar_arrive_expect_tx(bar(B_W1FULL + s1), S1_BYTES);
            tma_load_2d(sbase + OFF_W1 + s1 * S1_BYTES, &tmW1, kb * 64, c * HC, bar(B_W1FULL + s1));
            if (++s1 == NS1) { s1 = 0; ph1 ^= 1; }
          }
        // residual: when the last FC1 has retired (XEMPTY) LN(x) is dead -> reload the raw x tile into the same buffer
        mbar_wait(bar(B_XEMPTY), it & 1);
        mbar_arrive_expect_tx(bar(B_RESFULL), XN_BYTES);
        for (int kb = 0; kb < KB_X; ++kb) tma_load_2d(sbase + OFF_XN + kb * TM * 128, &tmX, kb * 64, tile * TM, bar(B_RESFULL));
      }
    }
  } else if (warp == 2) {
    // =============================== TMA producer: W2 K-blocks ===============================
    if (lane == 0) {
      int s2 = 0; uint32_t ph2 = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int c = 0; c < NCHUNK; ++c)
          for (int kb = 0; kb < KB_H; ++kb) {
            mbar_wait(bar(B_W2EMPTY + s2), ph2 ^ 1);
            mbar_arrive_expect_tx(bar(B_W2FULL + s2), S2_BYTES);
            tma_load_2d(sbase + OFF_W2 + s2 * S2_BYTES, &tmW2, c * HC + kb * 64, 0, bar(B_W2FULL + s2));
            if (++s2 == NS2) { s2 = 0; ph2 ^= 1; }
          }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(TM, HC, false, false);   // FC1: N = 128
      constexpr uint32_t idesc2 = umma_idesc_bf16(TM, D, false, false);    // FC2: N = 192
      int s1 = 0, s2 = 0; uint32_t ph1 = 0, ph2 = 0;
      uint32_t gf1 = 0, gf2 = 0;
      int it = 0;
      auto fc1 = [&]() {
        const uint32_t b = gf1 % NB, use = gf1 / NB;
        mbar_wait(bar(B_ACC1EMPTY + b), (use & 1) ^ 1);
        tc_fence_after();
        FMLP_STAMP(100 + 4 * (gf1 % NCHUNK));
        for (int kb = 0; kb < KB_X; ++kb) {
          mbar_wait(bar(B_W1FULL + s1), ph1);
          tc_fence_after();
          const uint32_t sa = sbase + OFF_XN + kb * TM * 128, sb = sbase + OFF_W1 + s1 * S1_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + b * HC, umma_desc_sw128(sa + k * 32, 16, 1024), umma_desc_sw128(sb + k * 32, 16, 1024), idesc1,
                      (kb | k) ? 1u : 0u);
          umma_commit(bar(B_W1EMPTY + s1));
          if (++s1 == NS1) { s1 = 0; ph1 ^= 1; }
        }
        umma_commit(bar(B_ACC1FULL + b));
        FMLP_STAMP(101 + 4 * (gf1 % NCHUNK));
        ++gf1;
      };
      auto fc2 = [&](int c) {
        const uint32_t b = gf2 % NB, use = gf2 / NB;
        mbar_wait(bar(B_HREADY + b), use & 1);
        FMLP_STAMP(102 + 4 * c);
        for (int kb = 0; kb < KB_H; ++kb) {
          mbar_wait(bar(B_W2FULL + s2), ph2);
          tc_fence_after();
          const uint32_t sa = sbase + OFF_H + b * H_BYTES + kb * TM * 128, sb = sbase + OFF_W2 + s2 * S2_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + ACC2_COL, umma_desc_sw128(sa + k * 32, 16, 1024), umma_desc_sw128(sb + k * 32, 16, 1024), idesc2,
                      (c | kb | k) ? 1u : 0u);
          umma_commit(bar(B_W2EMPTY + s2));
          if (++s2 == NS2) { s2 = 0; ph2 ^= 1; }
        }
        FMLP_STAMP(103 + 4 * c);
        umma_commit(bar(B_HEMPTY + b));
        ++gf2;
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(bar(B_XNREADY), it & 1);
        tc_fence_after();
        FMLP_STAMP(99);
        fc1();
        for (int c = 0; c < NCHUNK; ++c) {
          if (c + 1 < NCHUNK) {
            fc1();
            if (c + 2 == NCHUNK) umma_commit(bar(B_XEMPTY));            // last FC1 of the tile issued: x buffer frees when it retires
          }
          if (c == 0) { mbar_wait(bar(B_ACC2EMPTY), (it & 1) ^ 1); tc_fence_after(); }
          fc2(c);
        }
        umma_commit(bar(B_ACC2FULL));
      }
    }
  } else {
    // =============================== compute warps (3..18) ===============================
    const int cw = warp - 3;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int cq = cw >> 2;                    // column quarter (0..3)
    const int row = quad * 32 + lane;          // row within the tile
    const uint32_t tm_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(row & 7);
    uint32_t gc = 0;                           // global chunk counter
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const bool stamp = (warp == 3 && lane == 0);
      if (stamp) { bulk_wait_read0(); FMLP_STAMP(0); }      // previous tile's TMA store has finished reading the staging smem
      // ---- LayerNorm in place: thread (row, cq) owns columns [cq*48, cq*48+48) = 6 chunks of 8 bf16 ----
      mbar_wait(bar(B_XFULL), it & 1);
      if (stamp) FMLP_STAMP(1);
      uint4 v[6];
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int col = cq * 48 + i * 8, kb = col >> 6, ch = (col & 63) >> 3;
        v[i] = *reinterpret_cast<const uint4*>(sptr + OFF_XN + kb * TM * 128 + row * 128 + ((ch ^ sw) << 4));
        const uint32_t* w = &v[i].x;
#pragma unroll
        for (int j = 0; j < 4; ++j) { float a = bf16_lo(w[j]), b = bf16_hi(w[j]); s += a + b; q += a * a + b * b; }
      }
      ln_part[row * 4 + cq] = make_float2(s, q);
      asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
      for (int k = 1; k < 4; ++k) { const float2 o = ln_part[row * 4 + ((cq + k) & 3)]; s += o.x; q += o.y; }
      const float mean = s * (1.0f / D);
      const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int col = cq * 48 + i * 8, kb = col >> 6, ch = (col & 63) >> 3;
        uint32_t* w = &v[i].x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = col + 2 * j;
          float a = (bf16_lo(w[j]) - mean) * rstd * s_gamma[cc] + s_beta[cc];
          float b = (bf16_hi(w[j]) - mean) * rstd * s_gamma[cc + 1] + s_beta[cc + 1];
          w[j] = pack_bf16(a, b);
        }
        *reinterpret_cast<uint4*>(sptr + OFF_XN + kb * TM * 128 + row * 128 + ((ch ^ sw) << 4)) = v[i];
      }
      fence_proxy_async_smem();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_XNREADY));
      asm volatile("bar.sync 1, 512;" ::: "memory");   // ln_part reuse safety for the next tile
      if (stamp) FMLP_STAMP(2);

      // ---- hidden chunks: GELU epilogue into the FC2 A operand (thread: row, 32 of the chunk's 128 columns) ----
      for (int c = 0; c < NCHUNK; ++c, ++gc) {
        const uint32_t b = gc % NB, use = gc / NB;
        mbar_wait(bar(B_ACC1FULL + b), use & 1);
        tc_fence_after();
        if (stamp) FMLP_STAMP(10 + 4 * c);
        uint32_t r0[32];
        tmem_ld_32x32(tmem_base + tm_lane + b * HC + cq * 32, r0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_ACC1EMPTY + b));          // accumulator drained
        const float* bias = s_b1 + c * HC + cq * 32;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + 4 * j);
          pk[2 * j] = gelu_tanh_bf16x2(pack_bf16(__uint_as_float(r0[4 * j]) + bv.x, __uint_as_float(r0[4 * j + 1]) + bv.y));
          pk[2 * j + 1] = gelu_tanh_bf16x2(pack_bf16(__uint_as_float(r0[4 * j + 2]) + bv.z, __uint_as_float(r0[4 * j + 3]) + bv.w));
        }
        if (stamp) FMLP_STAMP(11 + 4 * c);
        mbar_wait(bar(B_HEMPTY + b), (use & 1) ^ 1);                // FC2 of the previous user has finished reading this H buffer
        if (stamp) FMLP_STAMP(12 + 4 * c);
        uint8_t* hrow = sptr + OFF_H + b * H_BYTES + (cq >> 1) * TM * 128 + row * 128;    // K-block (cq>>1), chunks (cq&1)*4 ..
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(hrow + ((((cq & 1) * 4 + ch) ^ sw) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_HREADY + b));
        if (stamp) FMLP_STAMP(13 + 4 * c);
      }
      // ---- final epilogue: out = acc2 + b2 + x  (residual tile in smem, result staged in the H buffers, TMA store) ----
      mbar_wait(bar(B_ACC2FULL), it & 1);
      mbar_wait(bar(B_RESFULL), it & 1);
      tc_fence_after();
      if (stamp) FMLP_STAMP(60);
      {
        const int col0 = cq * 48;
        uint32_t ra[32], rb[16];
        tmem_ld_32x32(tmem_base + tm_lane + ACC2_COL + col0, ra);
        tmem_ld_32x16(tmem_base + tm_lane + ACC2_COL + col0 + 32, rb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int col = col0 + 8 * j, kb = col >> 6, ch = (col & 63) >> 3;
          const uint32_t off = (uint32_t)(kb * TM * 128 + row * 128) + ((((uint32_t)ch) ^ sw) << 4);
          const uint4 xr = *reinterpret_cast<const uint4*>(sptr + OFF_XN + off);
          const uint32_t* xw = &xr.x;
          uint32_t ow[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = 8 * j + 2 * k;
            const float a0 = __uint_as_float(e < 32 ? ra[e & 31] : rb[(e - 32) & 15]);
            const float a1 = __uint_as_float(e + 1 < 32 ? ra[(e + 1) & 31] : rb[(e - 31) & 15]);
            ow[k] = pack_bf16(a0 + s_b2[col0 + e] + bf16_lo(xw[k]), a1 + s_b2[col0 + e + 1] + bf16_hi(xw[k]));
          }
          *reinterpret_cast<uint4*>(sptr + OFF_H + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_RESREAD));
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (stamp) {
        for (int kb = 0; kb < KB_X; ++kb) tma_store_2d(&tmOut, sbase + OFF_H + kb * TM * 128, kb * 64, tile * TM);
        bulk_commit();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_ACC2EMPTY));
      if (stamp) FMLP_STAMP(61);
    }
  }
  if (warp == 3 && lane == 0) bulk_wait0();                  // outstanding TMA stores complete before the CTA exits
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<fmlp::TMEM_COLS>(tmem_base);
  }
}

static long long* g_fmlp_dbg = nullptr;
void fused_mlp_set_debug(long long* buf) { g_fmlp_dbg = buf; }

bool fused_mlp_supported(int D, int hidden) { return D == fmlp::D && hidden == fmlp::HID; }

int launch_fused_mlp(cudaStream_t stream, const __nv_bfloat16* x, __nv_bfloat16* out, const float* gamma, const float* beta,
                     const __nv_bfloat16* w1, const float* b1, const __nv_bfloat16* w2, const float* b2, int M, int D, int hidden, float eps) {
  using namespace fmlp;
  if (D != fmlp::D || hidden != HID) { set_last_error("fused_mlp: only D=192, hidden=768"); return VITMARL_EINVAL; }
  if (M <= 0) return VITMARL_OK;
  CUtensorMap tmX, tmW1, tmW2, tmOut;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmX, x, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmOut, out, M, D, (uint64_t)D * 2, TM, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmW1, w1, HID, D, (uint64_t)D * 2, HC, 64))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmW2, w2, fmlp::D, HID, (uint64_t)HID * 2, fmlp::D, 64))) return rc;
  FusedMlpParams p{M, x, out, gamma, beta, b1, b2, eps, g_fmlp_dbg};
  cudaError_t e = cudaFuncSetAttribute(fused_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return check_cuda(e);
  const int tiles = (M + TM - 1) / TM;
  fused_mlp_kernel<<<min(tiles, num_sms()), THREADS, SMEM_BYTES, stream>>>(tmX, tmW1, tmW2, tmOut, p);
  return check_cuda(cudaGetLastError());
}

}  // namespace vitmarl
