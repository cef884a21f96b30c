// ViT encoder forward / backward orchestration (docs/VIT_SPEC.md) behind the C ABI.
// Drop-in for the flax `module.apply({'params': p}, x[B,H,W,C]) -> [B,D]` convention of the
// reference's vision encoder slot (gymnax_exchange/networks/vision_agent.py:16-24; the
// reference ships no ViT) and for its VJP inside the PPO loss
// (gymnax_exchange/jaxrl/MARL/ippo_rnn_JAXMARL.py:423-475).
//
// Every matrix product runs on the tcgen05 GEMM (gemm.cu): forward projections with fused
// bias / GELU / residual / pos-embed epilogues, dX products with the SAME weight buffer read
// as an MN-major operand (no transposed copies), dW products as MN-major x MN-major split-K
// GEMMs over the token dimension with fp32 red.add epilogues.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <new>

#include "common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

typedef __nv_bfloat16 bf16;

struct Dims {
  int B, H, W, C, P, D, L, heads, mlp, T, M, Kp;
  float eps;
};

static int get_dims(const VitmarlVitShape* s, Dims& d) {
  if (!s) return VITMARL_EINVAL;
  d.B = s->batch; d.H = s->img_h; d.W = s->img_w; d.C = s->channels; d.P = s->patch; d.D = s->dim; d.L = s->depth;
  d.heads = s->heads; d.mlp = s->mlp_dim; d.eps = s->ln_eps;
  if (d.B < 0 || d.P <= 0 || d.H % d.P || d.W % d.P || d.L < 0) { set_last_error("vit: bad shape"); return VITMARL_EINVAL; }
  d.T = (d.H / d.P) * (d.W / d.P);
  d.M = d.B * d.T;
  d.Kp = d.P * d.P * d.C;
  if (d.T != 64) { set_last_error("vit: (H/P)*(W/P) must be 64 tokens (attention kernel tile)"); return VITMARL_EINVAL; }
  if (d.D != d.heads * 64) { set_last_error("vit: head dim must be 64"); return VITMARL_EINVAL; }
  if (d.D % 64 || d.mlp % 64 || d.Kp % 64 || (d.P * d.C) % 8) { set_last_error("vit: D, mlp, P*P*C must be multiples of 64"); return VITMARL_EINVAL; }
  return VITMARL_OK;
}

// ---- packed parameter table (docs/VIT_SPEC.md "packed layout") ----------------------------------
enum { P_PE_W = 0, P_PE_B = 1, P_POS = 2, P_LAYER0 = 3, P_PER_LAYER = 12 };
enum { L_LN1_G = 0, L_LN1_B, L_QKV_W, L_QKV_B, L_OUT_W, L_OUT_B, L_LN2_G, L_LN2_B, L_FC1_W, L_FC1_B, L_FC2_W, L_FC2_B };
static inline int p_layer(int l, int k) { return P_LAYER0 + l * P_PER_LAYER + k; }
static inline int p_lnf_g(const Dims& d) { return P_LAYER0 + d.L * P_PER_LAYER; }
static inline int num_params(const Dims& d) { return p_lnf_g(d) + 2; }

static size_t param_elems(const Dims& d, int idx, bool* is_matrix) {
  *is_matrix = false;
  if (idx == P_PE_W) { *is_matrix = true; return (size_t)d.D * d.Kp; }
  if (idx == P_PE_B) return d.D;
  if (idx == P_POS) return (size_t)d.T * d.D;
  if (idx >= p_lnf_g(d)) return d.D;
  const int k = (idx - P_LAYER0) % P_PER_LAYER;
  switch (k) {
    case L_QKV_W: *is_matrix = true; return (size_t)3 * d.D * d.D;
    case L_QKV_B: return (size_t)3 * d.D;
    case L_OUT_W: *is_matrix = true; return (size_t)d.D * d.D;
    case L_FC1_W: *is_matrix = true; return (size_t)d.mlp * d.D;
    case L_FC1_B: return d.mlp;
    case L_FC2_W: *is_matrix = true; return (size_t)d.D * d.mlp;
    default: return d.D;
  }
}

// ---- workspace layout -------------------------------------------------------------------------
struct Ws {
  size_t patches, x, xm, ln1, ln2, qkv, att, hpre, hact, st1, st2, stf, dA, dB, dC, dH, dQKV, fold, pos_tile, total;
  size_t bxhat, bh2, bdh, dwf;          // fused MLP backward (training, D = 192): per-layer scratch reused by every layer
  size_t sz_md, sz_mh, sz_mq, sz_st;   // per-layer strides (bytes)
  size_t sz_fold, f_w1, f_b1, f_w2, f_wqkv, f_bq, f_bv, f_bo;   // inference: folded parameters per layer (vit_fold.cu), offsets within a layer's slot
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
// fused_train: the training path runs the MLP half of every block through the fused forward / backward kernels (no per-layer
// LN2 / FC1 / GELU activations are kept)
static bool train_fusable(const Dims& d) { return fused_mlp2_supported(d.D, d.mlp) && fused_mlp_bwd_supported(d.D, d.mlp) && d.L * 5 <= 64; }
static Ws layout(const Dims& d, bool save, bool fused_train = false) {
  Ws w{};
  const size_t M = (size_t)d.M;
  w.sz_md = al(M * d.D * 2); w.sz_mh = al(M * d.mlp * 2); w.sz_mq = al(M * 3 * d.D * 2); w.sz_st = al(M * 2 * 4);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
  const int nl = save ? d.L : 1;
  w.patches = take(M * d.Kp * 2);
  w.x = take(w.sz_md * (save ? d.L + 1 : 1));
  w.xm = save ? take(w.sz_md * nl) : w.x;          // inference: residual stream updated in place (out of place was measured: the
                                                   // per-thread residual re-reads cost both fused blocks ~15 %)
  w.ln1 = take(w.sz_md * nl);
  const bool mlp_saved = !(save && fused_train);
  w.ln2 = save ? (mlp_saved ? take(w.sz_md * nl) : 0) : w.ln1;
  w.qkv = take(w.sz_mq * nl);
  w.att = take(w.sz_md * nl);
  w.hact = mlp_saved ? take(w.sz_mh * nl) : 0;
  w.hpre = (save && mlp_saved) ? take(w.sz_mh * nl) : 0;
  w.st1 = save ? take(w.sz_st * nl) : 0;
  w.st2 = (save && mlp_saved) ? take(w.sz_st * nl) : 0;
  w.stf = save ? take(w.sz_st) : 0;
  if (save) {
    w.dA = take(w.sz_md); w.dB = take(w.sz_md); w.dC = take(w.sz_md);
    w.dQKV = take(w.sz_mq);
    if (mlp_saved) {
      w.dH = take(w.sz_mh);
    } else {
      w.bxhat = take(w.sz_md); w.bh2 = take(w.sz_mh); w.bdh = take(w.sz_mh);
      w.dwf = take((size_t)d.mlp * d.D * 4 + (size_t)d.mlp * 4);
    }
  }
  w.pos_tile = take((size_t)128 * d.D * 2);           // bf16 position-embedding tile (addend of the patch-embedding GEMM)
  if (!save || fused_train) {
    size_t o = 0;
    auto slot = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
    w.f_w1 = slot((size_t)d.mlp * d.D * 2); w.f_b1 = slot((size_t)d.mlp * 2); w.f_w2 = slot((size_t)d.D * d.mlp * 2);
    w.f_wqkv = slot((size_t)3 * d.D * d.D * 2); w.f_bq = slot((size_t)d.D * 2); w.f_bv = slot((size_t)d.D * 4); w.f_bo = slot((size_t)d.D * 4);
    w.sz_fold = o;
    w.fold = take(w.sz_fold * d.L);
  }
  w.total = off;
  return w;
}

// ---- optional per-launch timing (CUDA events on the launch stream), owned by a caller-created handle ---------------
enum { CAT_GEMM = 0, CAT_FUSED_MLP = 1, CAT_FUSED_ATTN = 2, CAT_ATTENTION = 3, CAT_LAYERNORM = 4, CAT_OTHER = 5, CAT_GEMM_DW = 6, CAT_GEMM_DX = 7, CAT_COUNT = 8 };
struct OpTiming {
  static constexpr int kMax = 16384;
  cudaEvent_t ev[2 * kMax];
  unsigned char cat[kMax];
  int created = 0;        // events created so far (lazily, two per logged launch)
  int n = 0;
  double flops = 0.0;     // algorithmic FLOPs of the tensor-core launches
};

// Everything a call may switch, resolved from VitmarlVitOptions (NULL = defaults).  No process-global mutable state: two
// threads / devices calling into the library concurrently never see each other's settings.
struct RunOpts {
  int fused = 1;                    // 0 unfused kernel sequence, 1 fused block kernels where the shape allows
  bool two_cta = true;
  FusedOpts fo;
  OpTiming* timing = nullptr;
  void* grads_flat = nullptr; size_t grads_flat_bytes = 0;
  bool accumulate = false;
  void* const* bucket_events = nullptr;
};

template <typename F>
static int timed(cudaStream_t st, const RunOpts& o, int cat, double flops, F&& launch) {
  OpTiming* t = o.timing;
  bool on = t && t->n < OpTiming::kMax;
  if (on && t->created < 2 * (t->n + 1)) {
    for (int i = t->created; i < 2 * (t->n + 1); ++i)
      if (cudaEventCreate(&t->ev[i]) != cudaSuccess) { on = false; break; } else t->created = i + 1;
  }
  if (on) cudaEventRecord(t->ev[2 * t->n], st);
  int rc = launch();
  // debugging aid: VITMARL_SYNC_EACH_LAUNCH=1 synchronises after every launch and names the kernel class that faulted
  static const bool sync_each = [] { const char* e = getenv("VITMARL_SYNC_EACH_LAUNCH"); return e && e[0] == '1'; }();
  if (sync_each && rc == VITMARL_OK) {
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      static const char* const names[CAT_COUNT] = {"gemm", "fused_mlp", "fused_attn", "attention", "layernorm", "other", "gemm_dW", "gemm_dX"};
      char msg[160];
      snprintf(msg, sizeof msg, "kernel class '%s' failed: %s", names[cat], cudaGetErrorString(e));
      set_last_error(msg);
      return VITMARL_ECUDA;
    }
  }
  if (on) {
    cudaEventRecord(t->ev[2 * t->n + 1], st);
    t->cat[t->n] = (unsigned char)cat;
    t->flops += flops;
    ++t->n;
  }
  return rc;
}

static int timed_gemm(cudaStream_t st, const RunOpts& o, GemmDesc& g) {
  g.allow_2cta = o.two_cta;
  return timed(st, o, CAT_GEMM, 2.0 * g.M * g.N * g.K, [&] { return launch_gemm(st, g); });
}

#define VM_TRY(x)            \
  do {                       \
    int rc__ = (x);          \
    if (rc__) return rc__;   \
  } while (0)

static GemmDesc gd(int M, int N, int K, const bf16* A, int lda, bool amn, const bf16* B, int ldb, bool bmn, void* C, int ldc, int epi) {
  GemmDesc g;
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.a_mn_major = amn; g.B = B; g.ldb = ldb; g.b_mn_major = bmn;
  g.C = C; g.ldc = ldc; g.epi = epi;
  return g;
}

static int vit_forward(cudaStream_t st, const RunOpts& o, const Dims& d, const void* const* prm, const bf16* x, float* y, uint8_t* ws, bool save,
                       bool reuse_folded = false, bool x_is_patches = false) {
  const bool fused_train = save && o.fused == 1 && train_fusable(d);
  const Ws w = layout(d, save, fused_train);
  const int M = d.M, D = d.D;
  auto PB = [&](int i) { return static_cast<const bf16*>(prm[i]); };
  auto PF = [&](int i) { return static_cast<const float*>(prm[i]); };
  const bf16* patches = (x_is_patches && !save) ? x : reinterpret_cast<const bf16*>(ws + w.patches);   // the caller may hand in the patch matrix itself
  if (x_is_patches && save) {   // training: the backward pass needs the patch matrix (patch-embedding dW) -> keep a copy in the workspace
    cudaError_t e = cudaMemcpyAsync(ws + w.patches, x, (size_t)M * d.Kp * sizeof(bf16), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return check_cuda(e);
  }
  auto X = [&](int l) { return reinterpret_cast<bf16*>(ws + w.x + (save ? (size_t)l * w.sz_md : 0)); };
  auto XM = [&](int l) { return reinterpret_cast<bf16*>(ws + w.xm + (save ? (size_t)l * w.sz_md : 0)); };
  auto LN1 = [&](int l) { return reinterpret_cast<bf16*>(ws + w.ln1 + (save ? (size_t)l * w.sz_md : 0)); };
  auto LN2 = [&](int l) { return reinterpret_cast<bf16*>(ws + w.ln2 + (save ? (size_t)l * w.sz_md : 0)); };
  auto QKV = [&](int l) { return reinterpret_cast<bf16*>(ws + w.qkv + (save ? (size_t)l * w.sz_mq : 0)); };
  auto ATT = [&](int l) { return reinterpret_cast<bf16*>(ws + w.att + (save ? (size_t)l * w.sz_md : 0)); };
  auto HACT = [&](int l) { return reinterpret_cast<bf16*>(ws + w.hact + (save ? (size_t)l * w.sz_mh : 0)); };
  auto HPRE = [&](int l) { return save ? reinterpret_cast<bf16*>(ws + w.hpre + (size_t)l * w.sz_mh) : nullptr; };
  auto ST1 = [&](int l) { return save ? reinterpret_cast<float*>(ws + w.st1 + (size_t)l * w.sz_st) : nullptr; };
  auto ST2 = [&](int l) { return save ? reinterpret_cast<float*>(ws + w.st2 + (size_t)l * w.sz_st) : nullptr; };

  // inference: fold LayerNorm scale/shift, the GELU 1/2, the softmax scale and the K / V biases into the projections
  // (vit_fold.cu), all layers in two launches (the second one needs the folded V bias of the first)
  const bool fuse_mlp2 = ((!save && o.fused == 1 && fused_mlp2_supported(D, d.mlp) && d.L * 5 <= 64) || fused_train);
  const bool fuse_attn2 = !save && o.fused == 1 && fused_attn2_supported(D, d.heads, d.T) && d.L * 5 <= 64;
  auto FW = [&](int l, size_t o) { return reinterpret_cast<bf16*>(ws + w.fold + (size_t)l * w.sz_fold + o); };
  auto FF = [&](int l, size_t o) { return reinterpret_cast<float*>(ws + w.fold + (size_t)l * w.sz_fold + o); };
  if ((fuse_mlp2 || fuse_attn2) && !reuse_folded) {
    FoldJobs jobs{}, jobs2{};
    for (int l = 0; l < d.L; ++l) {
      if (fuse_mlp2) {
        jobs.job[jobs.n++] = FoldJob{PB(p_layer(l, L_FC1_W)), PF(p_layer(l, L_FC1_B)), PF(p_layer(l, L_LN2_G)), PF(p_layer(l, L_LN2_B)),
                                     FW(l, w.f_w1), FW(l, w.f_b1), d.mlp, D, 1, 1.0f};
        jobs.job[jobs.n++] = FoldJob{PB(p_layer(l, L_FC2_W)), nullptr, nullptr, nullptr, FW(l, w.f_w2), nullptr, D, d.mlp, 0, 0.5f};
      }
      if (fuse_attn2) {
        const bf16* wqkv = PB(p_layer(l, L_QKV_W));
        const float* bqkv = PF(p_layer(l, L_QKV_B));
        const float *g1 = PF(p_layer(l, L_LN1_G)), *b1 = PF(p_layer(l, L_LN1_B));
        bf16* wf = FW(l, w.f_wqkv);
        const float qscale = 0.125f * 1.4426950408889634f;          // softmax(q.k / 8) evaluated as 2^(s - max)
        jobs.job[jobs.n++] = FoldJob{wqkv, bqkv, g1, b1, wf, FW(l, w.f_bq), D, D, 1, qscale};                                   // Q rows
        jobs.job[jobs.n++] = FoldJob{wqkv + (size_t)D * D, nullptr, g1, nullptr, wf + (size_t)D * D, nullptr, D, D, 0, 1.0f};   // K rows: bias cancels in the softmax
        jobs.job[jobs.n++] = FoldJob{wqkv + (size_t)2 * D * D, bqkv + 2 * D, g1, b1, wf + (size_t)2 * D * D, FF(l, w.f_bv), D, D, 0, 1.0f};   // V rows
        jobs2.job[jobs2.n++] = FoldJob{PB(p_layer(l, L_OUT_W)), PF(p_layer(l, L_OUT_B)), nullptr, FF(l, w.f_bv), nullptr, FF(l, w.f_bo), D, D, 0, 1.0f};
      }
    }
    VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_fold_params(st, jobs); }));
    if (jobs2.n) VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_fold_params(st, jobs2); }));
  }
  const bool pos_tiled = 128 % d.T == 0;
  bf16* pos_tile = reinterpret_cast<bf16*>(ws + w.pos_tile);
  if (pos_tiled && !reuse_folded) VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_pos_tile(st, PF(P_POS), pos_tile, d.T, D); }));
  // patch embedding: tokens = patches . Wpe^T + b + pos
  if (!x_is_patches)
    VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_patchify(st, x, reinterpret_cast<bf16*>(ws + w.patches), d.B, d.H, d.W, d.C, d.P); }));
  {
    GemmDesc g = gd(M, D, d.Kp, patches, d.Kp, false, PB(P_PE_W), d.Kp, false, X(0), D, EPI_STORE_BF16);
    g.bias = PF(P_PE_B); g.pos = PF(P_POS); g.pos_period = d.T;
    if (pos_tiled) g.pos_tile = pos_tile;
    VM_TRY(timed_gemm(st, o, g));
  }
  for (int l = 0; l < d.L; ++l) {
    if (fuse_attn2) {
      // inference: LN1 + QKV + softmax(QK^T)V + out-projection + residual in one kernel on the folded parameters
      VM_TRY(timed(st, o, CAT_FUSED_ATTN, 8.0 * M * D * D + 4.0 * M * d.T * D, [&] {
        return launch_fused_attn2(st, X(l), XM(l), FW(l, w.f_wqkv), FW(l, w.f_bq), PB(p_layer(l, L_OUT_W)), FF(l, w.f_bo), M, D, d.heads, d.eps, o.fo);
      }));
    } else {
    VM_TRY(timed(st, o, CAT_LAYERNORM, 0, [&] { return launch_layernorm(st, X(l), PF(p_layer(l, L_LN1_G)), PF(p_layer(l, L_LN1_B)), LN1(l), ST1(l), M, D, d.eps); }));
    {
      GemmDesc g = gd(M, 3 * D, D, LN1(l), D, false, PB(p_layer(l, L_QKV_W)), D, false, QKV(l), 3 * D, EPI_STORE_BF16);
      g.bias = PF(p_layer(l, L_QKV_B));
      VM_TRY(timed_gemm(st, o, g));
    }
    VM_TRY(timed(st, o, CAT_ATTENTION, 0, [&] { return launch_attention(st, QKV(l), ATT(l), d.B, d.heads); }));
    {
      GemmDesc g = gd(M, D, D, ATT(l), D, false, PB(p_layer(l, L_OUT_W)), D, false, XM(l), D, EPI_STORE_BF16);
      g.bias = PF(p_layer(l, L_OUT_B)); g.residual = X(l); g.ldr = D;
      VM_TRY(timed_gemm(st, o, g));
    }
    }
    if (fuse_mlp2) {
      // LN2 + FC1 + GELU + FC2 + residual in one CTA-pair kernel on the folded parameters (inference: in place; training: XM(l) ->
      // X(l+1) out of place -- the block input is all the fused backward kernel needs)
      VM_TRY(timed(st, o, CAT_FUSED_MLP, 4.0 * M * D * d.mlp, [&] {
        return launch_fused_mlp2(st, XM(l), X(l + 1), FW(l, w.f_w1), FW(l, w.f_b1), FW(l, w.f_w2), PF(p_layer(l, L_FC2_B)), M, D, d.mlp, d.eps, o.fo);
      }));
      continue;
    }
    VM_TRY(timed(st, o, CAT_LAYERNORM, 0, [&] { return launch_layernorm(st, XM(l), PF(p_layer(l, L_LN2_G)), PF(p_layer(l, L_LN2_B)), LN2(l), ST2(l), M, D, d.eps); }));
    {
      GemmDesc g = gd(M, d.mlp, D, LN2(l), D, false, PB(p_layer(l, L_FC1_W)), D, false, HACT(l), d.mlp, EPI_BIAS_GELU);
      g.bias = PF(p_layer(l, L_FC1_B)); g.C2 = HPRE(l);
      VM_TRY(timed_gemm(st, o, g));
    }
    {
      GemmDesc g = gd(M, D, d.mlp, HACT(l), d.mlp, false, PB(p_layer(l, L_FC2_W)), d.mlp, false, X(l + 1), D, EPI_STORE_BF16);
      g.bias = PF(p_layer(l, L_FC2_B)); g.residual = XM(l); g.ldr = D;
      VM_TRY(timed_gemm(st, o, g));
    }
  }
  float* stf = save ? reinterpret_cast<float*>(ws + w.stf) : nullptr;
  VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_final_ln_pool(st, X(d.L), PF(p_lnf_g(d)), PF(p_lnf_g(d) + 1), y, stf, d.B, d.T, D, d.eps); }));
  return VITMARL_OK;
}

static int vit_backward(cudaStream_t st, const RunOpts& o, const Dims& d, const void* const* prm, uint8_t* ws, const float* dy, void* const* dprm, bf16* dx) {
  const bool fused_train = o.fused == 1 && train_fusable(d);       // must match the forward pass that filled the workspace
  const Ws w = layout(d, true, fused_train);
  const int M = d.M, D = d.D, H = d.mlp;
  auto PB = [&](int i) { return static_cast<const bf16*>(prm[i]); };
  auto PF = [&](int i) { return static_cast<const float*>(prm[i]); };
  auto G = [&](int i) { return static_cast<float*>(dprm[i]); };
  bf16* patches = reinterpret_cast<bf16*>(ws + w.patches);
  auto X = [&](int l) { return reinterpret_cast<bf16*>(ws + w.x + (size_t)l * w.sz_md); };
  auto XM = [&](int l) { return reinterpret_cast<bf16*>(ws + w.xm + (size_t)l * w.sz_md); };
  auto LN1 = [&](int l) { return reinterpret_cast<bf16*>(ws + w.ln1 + (size_t)l * w.sz_md); };
  auto LN2 = [&](int l) { return reinterpret_cast<bf16*>(ws + w.ln2 + (size_t)l * w.sz_md); };
  auto QKV = [&](int l) { return reinterpret_cast<bf16*>(ws + w.qkv + (size_t)l * w.sz_mq); };
  auto ATT = [&](int l) { return reinterpret_cast<bf16*>(ws + w.att + (size_t)l * w.sz_md); };
  auto HACT = [&](int l) { return reinterpret_cast<bf16*>(ws + w.hact + (size_t)l * w.sz_mh); };
  auto HPRE = [&](int l) { return reinterpret_cast<bf16*>(ws + w.hpre + (size_t)l * w.sz_mh); };
  auto ST1 = [&](int l) { return reinterpret_cast<float*>(ws + w.st1 + (size_t)l * w.sz_st); };
  auto ST2 = [&](int l) { return reinterpret_cast<float*>(ws + w.st2 + (size_t)l * w.sz_st); };
  float* stf = reinterpret_cast<float*>(ws + w.stf);
  bf16* dA = reinterpret_cast<bf16*>(ws + w.dA);
  bf16* dB = reinterpret_cast<bf16*>(ws + w.dB);
  bf16* dC = reinterpret_cast<bf16*>(ws + w.dC);
  bf16* dH = reinterpret_cast<bf16*>(ws + w.dH);
  bf16* dQKV = reinterpret_cast<bf16*>(ws + w.dQKV);

  if (o.accumulate) {
    // micro-batch accumulation: the table keeps its contents, every gradient below is ADDED by fp32 reductions
  } else if (o.grads_flat) {
    // the caller carved the whole gradient table out of ONE buffer (GradAllReducer): a single memset instead of one per tensor
    cudaError_t e = cudaMemsetAsync(o.grads_flat, 0, o.grads_flat_bytes, st);
    if (e != cudaSuccess) return check_cuda(e);
  } else {
    for (int i = 0; i < num_params(d); ++i) {
      bool mat;
      size_t n = param_elems(d, i, &mat);
      cudaError_t e = cudaMemsetAsync(dprm[i], 0, n * sizeof(float), st);
      if (e != cudaSuccess) return check_cuda(e);
    }
  }
  // gradient buckets in the order the backward pass completes them: 0 = final LayerNorm, 1 + (L-1-l) = block l, L+1 = patch
  // embedding + position embedding.  An event per bucket lets the caller start that bucket's all-reduce on a side stream while
  // the next block's backward runs (jax.lax.pmean of the pmap trainer, ippo_rnn_JAXMARL_pmap.py:564-565).
  auto bucket_done = [&](int b) -> int {
    if (!o.bucket_events || !o.bucket_events[b]) return VITMARL_OK;
    return check_cuda(cudaEventRecord(static_cast<cudaEvent_t>(o.bucket_events[b]), st));
  };
  // dW[out,in] += dY^T . Xin   (both operands MN-major over the token dimension; split-K red.add)
  // db[out] += column sums of dY: handed to the same launch (the CTA-pair weight-gradient kernel folds it into its main loop as
  // one more MMA against a tile of ones, so dY is not read a second time; other shapes run the separate column-sum pass)
  auto dW = [&](float* dw, float* db, const bf16* dY, int n_out, const bf16* Xin, int n_in, float scale = 1.0f) {
    GemmDesc g = gd(n_out, n_in, M, dY, n_out, true, Xin, n_in, true, dw, n_in, EPI_ATOMIC_F32);
    g.colsum_a = db; g.allow_2cta = o.two_cta; g.out_scale = scale;
    return timed(st, o, CAT_GEMM_DW, 2.0 * g.M * g.N * g.K, [&] { return launch_gemm(st, g); });
  };
  // dXin[M,n_in] = dY[M,n_out] . W[n_out,n_in]   (W read as the MN-major B operand)
  auto dXg = [&](bf16* dXin, const bf16* dY, int n_out, const bf16* W, int n_in, int epi, const bf16* aux) {
    GemmDesc g = gd(M, n_in, n_out, dY, n_out, false, W, n_in, true, dXin, n_in, epi);
    g.residual = aux; g.ldr = n_in; g.allow_2cta = o.two_cta;
    return timed(st, o, CAT_GEMM_DX, 2.0 * g.M * g.N * g.K, [&] { return launch_gemm(st, g); });
  };

  VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_final_ln_pool_bwd(st, X(d.L), PF(p_lnf_g(d)), stf, dy, dA, G(p_lnf_g(d)), G(p_lnf_g(d) + 1), d.B, d.T, D); }));
  VM_TRY(bucket_done(0));
  for (int l = d.L - 1; l >= 0; --l) {
    // ---- MLP branch: x_{l+1} = xm + fc2(gelu(fc1(ln2(xm))))
    if (fused_train) {
      // one kernel recomputes the block from its input and produces dXM plus the operands of the two weight-gradient products
      bf16* bxhat = reinterpret_cast<bf16*>(ws + w.bxhat);
      bf16* bh2 = reinterpret_cast<bf16*>(ws + w.bh2);
      bf16* bdh = reinterpret_cast<bf16*>(ws + w.bdh);
      float* dwf = reinterpret_cast<float*>(ws + w.dwf);
      float* dbf = dwf + (size_t)H * D;
      auto FWl = [&](size_t off) { return reinterpret_cast<bf16*>(ws + w.fold + (size_t)l * w.sz_fold + off); };
      { cudaError_t e = cudaMemsetAsync(dwf, 0, ((size_t)H * D + H) * sizeof(float), st); if (e != cudaSuccess) return check_cuda(e); }
      // (the kernel also reduces two bias gradients while the tiles are on chip: db1' = column sums of dhpre, and the output-
      // projection bias gradient = column sums of dXM -- neither costs a pass over a [tokens, .] gradient any more)
      VM_TRY(timed(st, o, CAT_FUSED_MLP, 6.0 * M * D * H, [&] {
        return launch_fused_mlp_bwd(st, XM(l), dA, FWl(w.f_w1), FWl(w.f_b1), FWl(w.f_w2), bxhat, bh2, bdh, dB, M, D, H, d.eps, dbf,
                                    G(p_layer(l, L_OUT_B)), o.fo.dbg);
      }));
      VM_TRY(dW(G(p_layer(l, L_FC2_W)), G(p_layer(l, L_FC2_B)), dA, D, bh2, H, 0.5f));       // h2 = 2 gelu(hpre)
      VM_TRY(dW(dwf, nullptr, bdh, H, bxhat, D));                                              // gradient w.r.t. the FOLDED W1'
      VM_TRY(timed(st, o, CAT_OTHER, 0, [&] {
        return launch_unfold_grads(st, H, D, 1.0f, PB(p_layer(l, L_FC1_W)), PF(p_layer(l, L_LN2_G)), PF(p_layer(l, L_LN2_B)), dwf, dbf, G(p_layer(l, L_FC1_W)),
                                   G(p_layer(l, L_FC1_B)), G(p_layer(l, L_LN2_G)), G(p_layer(l, L_LN2_B)));
      }));
    } else {
    VM_TRY(dW(G(p_layer(l, L_FC2_W)), G(p_layer(l, L_FC2_B)), dA, D, HACT(l), H));
    VM_TRY(dXg(dH, dA, D, PB(p_layer(l, L_FC2_W)), H, EPI_MUL_GELU_GRAD, HPRE(l)));
    VM_TRY(dW(G(p_layer(l, L_FC1_W)), G(p_layer(l, L_FC1_B)), dH, H, LN2(l), D));
    VM_TRY(dXg(dC, dH, H, PB(p_layer(l, L_FC1_W)), D, EPI_STORE_BF16, nullptr));
    VM_TRY(timed(st, o, CAT_LAYERNORM, 0, [&] { return launch_layernorm_bwd(st, XM(l), PF(p_layer(l, L_LN2_G)), ST2(l), dC, dA, dB, G(p_layer(l, L_LN2_G)), G(p_layer(l, L_LN2_B)), M, D); }));
    }
    // ---- attention branch: xm = x_l + out(attn(qkv(ln1(x_l))))
    VM_TRY(dW(G(p_layer(l, L_OUT_W)), fused_train ? nullptr : G(p_layer(l, L_OUT_B)), dB, D, ATT(l), D));
    VM_TRY(dXg(dC, dB, D, PB(p_layer(l, L_OUT_W)), D, EPI_STORE_BF16, nullptr));
    // (the attention backward kernel reduces the QKV bias gradient -- column sums of dQKV -- from its gradient tiles)
    VM_TRY(timed(st, o, CAT_ATTENTION, 0, [&] { return launch_attention_bwd(st, QKV(l), dC, dQKV, d.B, d.heads, G(p_layer(l, L_QKV_B))); }));
    VM_TRY(dW(G(p_layer(l, L_QKV_W)), nullptr, dQKV, 3 * D, LN1(l), D));
    VM_TRY(dXg(dC, dQKV, 3 * D, PB(p_layer(l, L_QKV_W)), D, EPI_STORE_BF16, nullptr));
    VM_TRY(timed(st, o, CAT_LAYERNORM, 0, [&] { return launch_layernorm_bwd(st, X(l), PF(p_layer(l, L_LN1_G)), ST1(l), dC, dB, dA, G(p_layer(l, L_LN1_G)), G(p_layer(l, L_LN1_B)), M, D); }));
    VM_TRY(bucket_done(1 + (d.L - 1 - l)));
  }
  // ---- patch embedding
  VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_colsum(st, dA, G(P_POS), d.B, d.T * D); }));     // dpos[t,:] = sum over images
  VM_TRY(dW(G(P_PE_W), G(P_PE_B), dA, D, patches, d.Kp));
  VM_TRY(bucket_done(d.L + 1));
  if (dx) {
    VM_TRY(dXg(patches, dA, D, PB(P_PE_W), d.Kp, EPI_STORE_BF16, nullptr));   // reuse the patches buffer for d(patches)
    VM_TRY(timed(st, o, CAT_OTHER, 0, [&] { return launch_unpatchify(st, patches, dx, d.B, d.H, d.W, d.C, d.P); }));
  }
  return VITMARL_OK;
}

}  // namespace vitmarl

using namespace vitmarl;

extern "C" int vitmarl_vit_num_params(const VitmarlVitShape* s) {
  Dims d;
  if (get_dims(s, d)) return VITMARL_EINVAL;
  return num_params(d);
}

extern "C" long long vitmarl_vit_param_elems(const VitmarlVitShape* s, int index, int* is_bf16_matrix) {
  Dims d;
  if (get_dims(s, d) || index < 0 || index >= num_params(d)) return VITMARL_EINVAL;
  bool mat;
  size_t n = param_elems(d, index, &mat);
  if (is_bf16_matrix) *is_bf16_matrix = mat ? 1 : 0;
  return (long long)n;
}

extern "C" size_t vitmarl_vit_workspace_bytes(const VitmarlVitShape* s, int save_for_bwd) {
  Dims d;
  if (get_dims(s, d)) return 0;
  const bool save = save_for_bwd == 1;
  const size_t a = layout(d, save, false).total, b = layout(d, save, save && train_fusable(d)).total;
  return a > b ? a : b;       // either training layout (fused / unfused MLP half, VitmarlVitOptions::fused) fits
}

static RunOpts resolve(const VitmarlVitOptions* v) {
  RunOpts o;
  if (!v) return o;
  if (v->fused >= 0) o.fused = v->fused ? 1 : 0;
  if (v->gemm_2cta >= 0) o.two_cta = v->gemm_2cta != 0;
  if (v->pdl >= 0) o.fo.pdl = v->pdl != 0;
  if (v->attn_flags >= 0) o.fo.attn_flags = v->attn_flags & 0xff;
  o.fo.dbg = v->debug_timeline;
  o.timing = static_cast<OpTiming*>(v->timing);
  o.grads_flat = v->grads_flat; o.grads_flat_bytes = v->grads_flat_bytes;
  o.accumulate = v->accumulate > 0;
  o.bucket_events = v->bucket_events;
  return o;
}

extern "C" int vitmarl_vit_fwd_ex(void* stream, const VitmarlVitShape* s, const void* const* params, const void* x, float* y,
                                  void* workspace, size_t workspace_bytes, int save_for_bwd, const VitmarlVitOptions* opt) {
  Dims d;
  VM_TRY(get_dims(s, d));
  if (d.B == 0) return VITMARL_OK;
  if (!params || !x || !y || !workspace) return VITMARL_EINVAL;
  const bool x_is_patches = (save_for_bwd & VITMARL_VIT_INPUT_PATCHES) != 0;
  save_for_bwd &= ~VITMARL_VIT_INPUT_PATCHES;
  const bool save = save_for_bwd == 1;
  if (x_is_patches && (reinterpret_cast<uintptr_t>(x) & 15)) { set_last_error("vit_fwd: patch-matrix input must be 16-byte aligned"); return VITMARL_EINVAL; }
  if (workspace_bytes < vitmarl_vit_workspace_bytes(s, save ? 1 : 0)) { set_last_error("vit_fwd: workspace too small"); return VITMARL_EINVAL; }
  return vit_forward(static_cast<cudaStream_t>(stream), resolve(opt), d, params, static_cast<const bf16*>(x), y, static_cast<uint8_t*>(workspace), save,
                     save_for_bwd == 2, x_is_patches);
}

extern "C" int vitmarl_vit_fwd(void* stream, const VitmarlVitShape* s, const void* const* params, const void* x, float* y,
                               void* workspace, size_t workspace_bytes, int save_for_bwd) {
  return vitmarl_vit_fwd_ex(stream, s, params, x, y, workspace, workspace_bytes, save_for_bwd, nullptr);
}

extern "C" int vitmarl_vit_bwd_ex(void* stream, const VitmarlVitShape* s, const void* const* params, void* workspace,
                                  size_t workspace_bytes, const float* dy, void* const* dparams, void* dx, const VitmarlVitOptions* opt) {
  Dims d;
  VM_TRY(get_dims(s, d));
  if (d.B == 0) return VITMARL_OK;
  if (!params || !workspace || !dy || !dparams) return VITMARL_EINVAL;
  if (workspace_bytes < vitmarl_vit_workspace_bytes(s, 1)) { set_last_error("vit_bwd: workspace too small"); return VITMARL_EINVAL; }
  return vit_backward(static_cast<cudaStream_t>(stream), resolve(opt), d, params, static_cast<uint8_t*>(workspace), dy, dparams, static_cast<bf16*>(dx));
}

extern "C" int vitmarl_vit_bwd(void* stream, const VitmarlVitShape* s, const void* const* params, void* workspace,
                               size_t workspace_bytes, const float* dy, void* const* dparams, void* dx) {
  return vitmarl_vit_bwd_ex(stream, s, params, workspace, workspace_bytes, dy, dparams, dx, nullptr);
}

extern "C" int vitmarl_vit_num_buckets(const VitmarlVitShape* s) {
  Dims d;
  if (get_dims(s, d)) return VITMARL_EINVAL;
  return d.L + 2;
}

// ---- timing handle: CUDA-event log of the launches issued by calls that carry it in their options --------------------------
extern "C" void* vitmarl_timing_create(void) { return new (std::nothrow) OpTiming(); }

extern "C" void vitmarl_timing_destroy(void* h) {
  OpTiming* t = static_cast<OpTiming*>(h);
  if (!t) return;
  for (int i = 0; i < t->created; ++i) cudaEventDestroy(t->ev[i]);
  delete t;
}

extern "C" int vitmarl_timing_reset(void* h) {
  OpTiming* t = static_cast<OpTiming*>(h);
  if (!t) return VITMARL_EINVAL;
  t->n = 0; t->flops = 0.0;
  return VITMARL_OK;
}

// Per-category totals: ms8 / n8 are arrays of 8 (0 gemm, 1 fused mlp, 2 fused attention block, 3 attention,
// 4 layernorm, 5 other, 6 dW gemm, 7 dX gemm); flops = algorithmic FLOPs of the tensor-core launches (categories 0-2, 6, 7).
// Synchronises on the last logged launch.
extern "C" int vitmarl_timing_read(void* h, double* ms8, long long* n8, double* flops) {
  OpTiming* t = static_cast<OpTiming*>(h);
  if (!t) return VITMARL_EINVAL;
  for (int i = 0; i < CAT_COUNT; ++i) { if (ms8) ms8[i] = 0.0; if (n8) n8[i] = 0; }
  if (t->n > 0) {
    cudaError_t e = cudaEventSynchronize(t->ev[2 * t->n - 1]);
    if (e != cudaSuccess) return check_cuda(e);
    for (int i = 0; i < t->n; ++i) {
      float x = 0.f;
      if (cudaEventElapsedTime(&x, t->ev[2 * i], t->ev[2 * i + 1]) == cudaSuccess && ms8) ms8[t->cat[i]] += x;
      if (n8) n8[t->cat[i]] += 1;
    }
  }
  if (flops) *flops = t->flops;
  return VITMARL_OK;
}

// Test hook: C[M,N] (fp32, accumulated into) += A^T . B and colsum[M] += column sums of A, A [K,M] / B [K,N] bf16 row-major
// (the dW + bias-gradient launch of the backward pass).
extern "C" int vitmarl_debug_gemm_dw(void* stream, int M, int N, int K, const void* A, const void* B, float* C, float* colsum, int flags) {
  GemmDesc g;
  g.M = M; g.N = N; g.K = K;
  g.A = static_cast<const bf16*>(A); g.lda = M; g.a_mn_major = true;
  g.B = static_cast<const bf16*>(B); g.ldb = N; g.b_mn_major = true;
  g.C = C; g.ldc = N; g.epi = EPI_ATOMIC_F32; g.colsum_a = colsum; g.allow_2cta = !(flags & VITMARL_GEMM_NO_2CTA);
  return launch_gemm(static_cast<cudaStream_t>(stream), g);
}
