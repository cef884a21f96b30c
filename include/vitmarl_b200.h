/* vitmarl_b200 -- C ABI of the B200-native rollout-and-encode hot path.
 *
 * Drop-in boundary for the three reference call sites of hiepday3324/ViT-MARL (the
 * reference is pure Python/JAX; it has no FFI of its own, so these entry points are what
 * a jax.ffi / XLA custom-call binding -- or the ctypes binding in vitmarl_b200/_capi.py --
 * binds; INTEGRATION.md shows the reference-side stubs):
 *
 *   order-book step  gymnax_exchange/jaxob/JaxOrderBookArrays.py:720-752
 *                    scan_through_entire_array_save_bidask (called under jax.vmap from
 *                    gymnax_exchange/jaxen/marl_env.py:377-384), :665-685 scan_through_entire_array
 *   read-outs        JaxOrderBookArrays.py:881-898 get_best_bid_and_ask_inclQuants,
 *                    :1075-1106 get_L2_state, :1108-1140 get_vision_L2_state
 *   observation      gymnax_exchange/jaxen/vision_env.py:2709-2721 _get_obs_vision,
 *                    :2804-2854 normalize_vision_obs; marl_env.py:392-393,467,685-711
 *                    (_ffill_best_prices + mid price)
 *   encoder          flax `module.apply({'params': p}, x[B,H,W,C])` convention of
 *                    gymnax_exchange/networks/vision_agent.py:16-24 (the ViT itself is
 *                    builder-specified: docs/VIT_SPEC.md) and its VJP used by
 *                    jaxrl/MARL/ippo_rnn_JAXMARL.py:423-475.
 *
 * Conventions: every pointer is caller-owned DEVICE memory unless the function name ends
 * in _host; no allocation, no synchronisation, work is only enqueued on `stream`
 * (a cudaStream_t passed as void*).  Returns 0 or a negative VITMARL_E* code; never aborts.
 * Threading: no process-global mutable state -- tuning / debug / measurement switches are per-call
 * (VitmarlVitOptions, `flags` arguments, caller-owned timing handles); the last-error text is thread-local.
 * All book / message / trade arrays are int32, row-major, layouts of
 * gymnax_exchange/jaxob/jaxob_constants.py:36-52,76-83.
 */
#ifndef VITMARL_B200_H_
#define VITMARL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITMARL_OK 0
#define VITMARL_EINVAL (-1)       /* bad shape / null pointer / misaligned buffer           */
#define VITMARL_EUNSUPPORTED (-2) /* cancel_mode 2/3 (PRNG-dependent), simulator_mode 1      */
#define VITMARL_ECUDA (-3)        /* launch failure (cudaGetLastError)                      */
#define VITMARL_ENODEVICE (-4)    /* no sm_100 device                                       */

#define VITMARL_ABI_VERSION 3

int vitmarl_abi_version(void);
/* Human-readable text for the last CUDA error seen by this thread (static storage). */
const char* vitmarl_last_error(void);

/* ---- stage 1: order-book step ------------------------------------------------------ */

/* Replaces job.scan_through_entire_array_save_bidask under vmap (JaxOrderBookArrays.py:720-752):
 * for each of E environments process M messages [type, side, qty, price, oid, tid, t_s, t_ns]
 * in order through add / cancel / match (lines 62-330, 356-478, 617-661) and record the best
 * ask / best bid [price, volume at price] after each of the LAST n_keep messages (lines 881-898).
 *   asks_in/bids_in  [E,N,6]  (may alias asks_out/bids_out: in-place update)
 *   trades_in        [E,T,8]  or NULL -> trades start at -1 (marl_env.py:377)
 *   msgs             [E,M,8]
 *   trades_out       [E,T,8]
 *   best_asks/bids   [E,n_keep,2] or NULL (scan_through_entire_array, lines 665-685)
 *   cancel_mode      JAXLOB_Configuration.cancel_mode (0 or 1; 2/3 -> VITMARL_EUNSUPPORTED)
 *   init_id          JAXLOB_Configuration.init_id (-2)
 * Limits: 1 <= N <= 256, 1 <= T <= 1024, M >= 0, 0 <= n_keep. */
int vitmarl_lob_step(void* stream, int E, int N, int T, int M, int n_keep,
                     const int32_t* asks_in, const int32_t* bids_in, const int32_t* trades_in,
                     const int32_t* msgs,
                     int32_t* asks_out, int32_t* bids_out, int32_t* trades_out,
                     int32_t* best_asks, int32_t* best_bids,
                     int cancel_mode, int32_t init_id);

/* Replaces job.get_best_bid_and_ask_inclQuants under vmap (JaxOrderBookArrays.py:881-898).
 *   best_ask/best_bid [E,2] */
int vitmarl_lob_best_bid_ask(void* stream, int E, int N, const int32_t* asks, const int32_t* bids,
                             int32_t* best_ask, int32_t* best_bid);

/* ---- stage 2: observation render --------------------------------------------------- */

#define VITMARL_IMG_NONE 0
#define VITMARL_IMG_U8 1   /* uint8 {0,1}          */
#define VITMARL_IMG_BF16 2 /* bfloat16 {0.0,1.0}   */
/* bfloat16 raster written directly as the ViT's patch matrix [E, (H/p)*(W/p), p*p*2] (token = py*(W/p)+px, feature =
 * (ph, pw, c); docs/VIT_SPEC.md) -- the same values as VITMARL_IMG_BF16 in the order vitmarl_vit_fwd's patch-embedding GEMM
 * reads them (pass save_for_bwd | VITMARL_VIT_INPUT_PATCHES there).  p % 4 == 0, H % p == 0, W % p == 0. */
#define VITMARL_IMG_BF16_PATCHES 3
#define VITMARL_IMG_BF16_PATCHES_OF(p) (((p) << 8) | VITMARL_IMG_BF16_PATCHES)

/* Replaces job.get_vision_L2_state (JaxOrderBookArrays.py:1108-1140) and, when `norm` and
 * `mid_price` are given, ExecutionAgent.normalize_vision_obs (vision_env.py:2804-2854);
 * optionally also rasterises the book (docs/RENDER_SPEC.md; builder-defined).
 *   raw        int32 [E,n_levels,2,2]  (level, (price, volume), (ask, bid))      or NULL
 *   l2         int32 [E,4*n_levels]    get_L2_state layout (lines 1075-1106)      or NULL
 *   mid_price  float [E]               WorldState.mid_price                       or NULL
 *   norm       float [E,n_levels,3,2]  (level, (gap, log1p vol, log1p cum), (ask,bid)) or NULL
 *   image      [E,H,W,2] of img_dtype  (channel 0 ask, 1 bid)                     or NULL
 * Limits: n_levels <= 32, W % 8 == 0, 2*H <= 12*N. */
int vitmarl_lob_render(void* stream, int E, int N, int n_levels, int tick_size,
                       const int32_t* asks, const int32_t* bids, const float* mid_price,
                       int32_t* raw, int32_t* l2, float* norm,
                       void* image, int img_dtype, int H, int W);

/* Fused environment step = vitmarl_lob_step (trades start at -1, n_keep = M)
 *   + _ffill_best_prices on both best-price tracks and the new mid price
 *     (marl_env.py:392-393, 466-467, 685-711)
 *   + vitmarl_lob_render on the updated book with that mid price
 * in ONE kernel: the book is read once and written once per step.
 *   last_ask_price/last_bid_price [E]   previous step's final best prices
 *                                       (state.world_state.best_asks[-1,0] / best_bids[-1,0])
 *   best_asks/best_bids [E,M,2]         forward-filled
 *   mid_price [E] float                 (ffilled bid[-1] + ask[-1]) / 2
 * raw / l2 / norm / image as in vitmarl_lob_render (each nullable). Requires M >= 1. */
int vitmarl_env_step(void* stream, int E, int N, int T, int M,
                     const int32_t* asks_in, const int32_t* bids_in, const int32_t* msgs,
                     const int32_t* last_ask_price, const int32_t* last_bid_price,
                     int32_t* asks_out, int32_t* bids_out, int32_t* trades_out,
                     int32_t* best_asks, int32_t* best_bids, float* mid_price,
                     int n_levels, int tick_size, int32_t* raw, int32_t* l2, float* norm,
                     void* image, int img_dtype, int H, int W,
                     int cancel_mode, int32_t init_id);

/* vitmarl_env_step with named arguments plus what the callers either side of it need:
 *   last_price_stride   element stride of last_ask_price / last_bid_price: 1 = dense [E]; 2*M = the caller points at
 *                       best_asks[0, M-1, 0] / best_bids[0, M-1, 0] of the PREVIOUS step (state.world_state.best_asks[-1, 0],
 *                       marl_env.py:392-393).  Those buffers may be the SAME ones passed as best_asks / best_bids outputs: each
 *                       environment's last price is read before its rows are rewritten.
 *   n_stat_agents (0..4), stat_agent_ids, trade_stats [E, n_stat_agents, 8]
 *                       the integer trade reductions of the reward functions for these trader ids, computed from the step's
 *                       trade log while it is still on chip; row layout and arithmetic of vitmarl_agent_trade_stats
 *                       (get_agent_trades JaxOrderBookArrays.py:824-831; vision_env.py:2076-2078,2156-2163,2191; mm_env.py:1906-1936).
 *   trades_out          may be NULL when n_stat_agents > 0: the [T,8] log is then never written to HBM (-3.2 KB per env-step).
 *   time_in [E,2], time_out [E,2], delta_time [E]   (all or none; ABI 3)
 *                       the world clock of MARLEnv.step_env (marl_env.py:406,468,482): time_out = final_time = the (seconds, ns)
 *                       columns of the step's LAST message, delta_time = float32 of
 *                       ((final_time[0] + final_time[1]/1e9) - time_in[0]) - time_in[1]/1e9, every operation rounded to float32 as
 *                       JAX does with x64 disabled.  time_out may alias time_in.  Not written when M == 0. */
typedef struct VitmarlEnvStepArgs {
  int E, N, T, M;
  const int32_t* asks_in; const int32_t* bids_in; const int32_t* msgs;
  const int32_t* last_ask_price; const int32_t* last_bid_price; int last_price_stride;
  int32_t* asks_out; int32_t* bids_out; int32_t* trades_out;
  int32_t* best_asks; int32_t* best_bids; float* mid_price;
  int n_levels; int tick_size; int32_t* raw; int32_t* l2; float* norm;
  void* image; int img_dtype; int H; int W;
  int cancel_mode; int32_t init_id;
  int n_stat_agents; int32_t stat_agent_ids[4]; int32_t* trade_stats;
  const int32_t* time_in; int32_t* time_out; float* delta_time;
} VitmarlEnvStepArgs;
int vitmarl_env_step2(void* stream, const VitmarlEnvStepArgs* args);

/* ---- callers either side of the order-book step --------------------- */

/* Replaces job.getCancelMsgs under vmap (JaxOrderBookArrays.py:756-782): for each env the first `size` rows
 * (ascending) of `bookside` [E,N,6] whose trader id equals agent_id become cancel messages
 * [2, side, qty, price, oid, tid, t_s, t_ns]; missing ones are [2, side, 0, 0, 0, 0, t_s, t_ns].
 *   cancel_time [E,2], out [E,size,8] */
int vitmarl_get_cancel_msgs(void* stream, int E, int N, int size, const int32_t* bookside, int agent_id, int side,
                            const int32_t* cancel_time, int32_t* out);

/* Auto-reset of MARLEnv.step (marl_env.py:737-766) for the world-state leaves this library owns: where done[e] != 0, env e's
 * books / trades are replaced by those of its data window window_index[e] (base_env.py:215-231, index_tree(init_states_array)),
 * best_asks / best_bids [E,M,2] are tiled from the window's initial best ask / bid [n_windows,2] (marl_env.py:186-189) and
 * mid_price = float32((best_bid + best_ask) / 2) (:190).  Environments that are not done are untouched (the reference builds a
 * full reset state for every env and selects).  window_index is the caller's jax.random.randint draw (base_env.py:219-222);
 * init_trades [n_windows,T,8] may be NULL (= all -1).  All buffers int32 device memory, 16-byte aligned trades.
 * A done environment whose window_index is outside [0, n_windows) is left untouched and flagged: if `bad_window` (int32 [1],
 * device, nullable) is given it receives 1 (JAX would clamp the gather; silently reading out of bounds is not an option here). */
int vitmarl_auto_reset(void* stream, int E, int N, int T, int M, int n_windows, const int32_t* done, const int32_t* window_index,
                       const int32_t* init_asks, const int32_t* init_bids, const int32_t* init_trades,
                       const int32_t* init_best_asks, const int32_t* init_best_bids, int32_t* asks, int32_t* bids,
                       int32_t* trades, int32_t* best_asks, int32_t* best_bids, float* mid_price, int32_t* bad_window);

/* Replaces job.get_agent_trades under vmap (JaxOrderBookArrays.py:824-831): rows of trades [E,T,8] that are executed
 * (price >= 0) and involve agent_id as passive or aggressive trader are kept, all others zeroed. */
int vitmarl_get_agent_trades(void* stream, int E, int T, const int32_t* trades, int agent_id, int32_t* out);

/* Replaces `_filter_messages` under vmap (vision_env.py:622-684; the same text in mm_env.py:509-571 and exec_env.py:611-673):
 * cancellations are netted against new orders at the same non-zero price -- the t-th matching action row is paired with the
 * t-th matching cancel row, both lose rel = (c_qty >= a_qty) * a_qty, and action rows left with quantity 0 become all-zero
 * dummy messages.  action_msgs / cnl_msgs / outputs [E,n,8] int32, n <= 32, 16-byte aligned; outputs may alias the inputs. */
int vitmarl_filter_messages(void* stream, int E, int n, const int32_t* action_msgs, const int32_t* cnl_msgs,
                            int32_t* action_out, int32_t* cnl_out);

/* The per-step trade reductions of the reward functions, fused in one pass over trades [E,T,8] (int32 wrap-around
 * arithmetic, as XLA): out [E,8] = [sum qty, sum |qty|, c_rl, buyQuant, sellQuant, TradedVolume, inventory_delta,
 * sum |qty| of the other executed trades] of the rows get_agent_trades keeps for agent_id, where
 *   sum qty        = agent_trades[:,1].sum()                                  (vision_env.py:2076-2077, before jnp.abs)
 *   sum |qty|      = jnp.abs(agentTrades[:,1]).sum()                          (vision_env.py:2163)
 *   c_rl           = (agentTrades[:,0] // tick_size * |agentTrades[:,1]|).sum() (vision_env.py:2191)
 *   buy/sellQuant, TradedVolume = buy + sell, inventory_delta = buy - sell     (mm_env.py:1917-1933)
 * trades and out 16-byte aligned device memory. */
int vitmarl_agent_trade_stats(void* stream, int E, int T, const int32_t* trades, int agent_id, int tick_size, int32_t* out);

/* Default action -> message tables of the two agents, one call for all E environments:
 *   ExecutionAgent._getActionMsgs_fixedQuant_complex (vision_env.py:1046-1142; action_space "fixed_quants_complex", the
 *   Execution_EnvironmentConfig default jaxob_config.py:108)  -> out [E,4,8]: four limit orders at FT / M / NT / PP
 *   MarketMakingAgent._getActionMsgs_spread_skew (mm_env.py:1352-1491; action_space "spread_skew", the MarketMaking_EnvironmentConfig
 *   default jaxob_config.py:34; multiplier_type 0 = "tick" (default), 1 = "spread")  -> out [E,2,8]: a bid and an ask
 * best_*_price + price_stride: world_state.best_asks[-1][0] / best_bids[-1][0] read by stride from the best-price tracks (as in
 * vitmarl_env_step2); action, is_sell_task, task_to_execute, quant_executed: int32 [E]; time: int32 [E,2] (world_state.time; the
 * reference adds time_delay_obs_act to BOTH fields).  Integer arithmetic wraps like XLA's; float arithmetic is float32 with one
 * rounding per operation and jax.numpy.floor_divide's algorithm, float -> int32 by truncation.  Parity: bit-exact vs the NumPy
 * restatement oracle/action_oracle.py; unpinned against JAX itself (not installed in this image). */
int vitmarl_exec_action_msgs_fixed_quants_complex(void* stream, int E, const int32_t* action, const int32_t* best_ask_price,
                                                  const int32_t* best_bid_price, int price_stride, const int32_t* is_sell_task,
                                                  const int32_t* task_to_execute, const int32_t* quant_executed, const int32_t* time,
                                                  int trader_id, int tick_size, int n_ticks_in_book, int fixed_quant_value,
                                                  int time_delay_obs_act, int placeholder_order_id, int32_t* out);
int vitmarl_mm_action_msgs_spread_skew(void* stream, int E, const int32_t* action, const int32_t* best_ask_price,
                                       const int32_t* best_bid_price, int price_stride, const int32_t* time, int trader_id, int tick_size,
                                       float spread_multiplier, float skew_multiplier, int multiplier_type, int fixed_quant_value,
                                       int time_delay_obs_act, int placeholder_order_id, int32_t* out);

/* Message assembly of MARLEnv.step_env (marl_env.py:272-344): combined[e] = [cancel_msgs[e]; action_msgs[e] with
 * order ids renumbered to order_id_counter[e] - arange(Ma) and rows permuted by perm[e] (jax.random.permutation
 * indices computed by the caller; NULL = no shuffle); data messages] where the data messages are
 * lax.dynamic_slice(message_data [n_total,8], start_index[e] + n_data*step_counter[e], n_data) (start clamped,
 * base_env.py:341-371) with the first six fields zeroed when time_s >= end_time_s[e] (fixed_time episodes; NULL =
 * fixed_steps).  combined [E, Mc+Ma+n_data, 8]; new_order_id_counter [E] = counter - Ma (nullable). */
int vitmarl_build_step_msgs(void* stream, int E, int n_total, int n_data, int Mc, int Ma, const int32_t* message_data,
                            const int32_t* start_index, const int32_t* step_counter, const int32_t* end_time_s,
                            const int32_t* cancel_msgs, const int32_t* action_msgs, const int32_t* perm,
                            const int32_t* order_id_counter, int32_t* combined, int32_t* new_order_id_counter);

/* ---- stage 3: ViT encoder ---------------------------------------------------------- */

#define VITMARL_EPI_STORE_BF16 0 /* C(bf16) = acc (+bias) (+pos[row % period]) (+residual)        */
#define VITMARL_EPI_BIAS_GELU 1  /* C(bf16) = gelu_tanh(acc + bias)                              */
#define VITMARL_EPI_STORE_F32 2  /* C(fp32) = acc (+bias)                                        */
#define VITMARL_EPI_ATOMIC_F32 3 /* C(fp32) += out_scale * acc  (split-K, C pre-initialised)     */
#define VITMARL_EPI_MUL_GELU_GRAD 4 /* C(bf16) = acc * gelu_tanh'(residual[row,col])             */

/* The tensor-core building block of the encoder (TMA -> tcgen05.mma -> TMEM epilogue):
 *   C[M,N] = epi( A[M,K] . B[N,K]^T ),  A/B bf16, fp32 accumulate.
 * Operand X is K-major when element (r,k) is X[r*ldx + k] (activations [tokens,features],
 * flax Dense kernels stored [out,in]) and MN-major when it is X[k*ldx + r].
 * Limits: N % 64 == 0, K % 8 == 0, lda/ldb % 8 == 0, 16-byte aligned bases.
 * flags: VITMARL_GEMM_NO_2CTA = use the 1-CTA kernels only (default: the CTA-pair tcgen05 cta_group::2 kernels where they apply). */
#define VITMARL_GEMM_NO_2CTA 1
int vitmarl_gemm_bf16(void* stream, int M, int N, int K,
                      const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major,
                      void* C, int ldc, int epi, const float* bias, const void* residual, int ldr,
                      const float* pos, int pos_period, float out_scale, int flags);

/* Test hook for the weight-gradient launch of the backward pass: C[M,N] (fp32) += A^T . B, colsum[M] += column sums of A
 * (A [K,M], B [K,N] bf16 row-major; colsum may be NULL); flags as vitmarl_gemm_bf16. */
int vitmarl_debug_gemm_dw(void* stream, int M, int N, int K, const void* A, const void* B, float* C, float* colsum, int flags);

/* The attention core of the unfused encoder path on its own (test / measurement hook; TMA + tcgen05 + TMEM, csrc/attention_tc.cu):
 * 64 tokens per image, head dim 64.  qkv [B*64, 3*64*heads] bf16 (q | k | v, head-major), out / dout [B*64, 64*heads],
 * dqkv like qkv.  forward: out = softmax(q k^T / 8) v per (image, head); backward: dqkv from (qkv, dout). */
int vitmarl_attention_fwd(void* stream, int B, int heads, const void* qkv, void* out);
int vitmarl_attention_bwd(void* stream, int B, int heads, const void* qkv, const void* dout, void* dqkv);

/* Shape of the encoder (docs/VIT_SPEC.md): pre-LN ViT, learned position embedding, no class
 * token, final LayerNorm then mean pool -> [B, dim].  Requires (H/P)*(W/P) == 64 tokens and
 * head dim 64 (ViT-Tiny/8 @ 64x64: dim 192, heads 3; ViT-S/16 @ 128x128: dim 384, heads 6). */
typedef struct VitmarlVitShape {
  int batch;      /* B images                                      */
  int img_h;      /* H                                             */
  int img_w;      /* W                                             */
  int channels;   /* C (2: ask, bid)                               */
  int patch;      /* P                                             */
  int dim;        /* D                                             */
  int depth;      /* L                                             */
  int heads;      /* h = D / 64                                    */
  int mlp_dim;    /* 4 * D                                         */
  float ln_eps;   /* 1e-6 (flax LayerNorm default)                 */
} VitmarlVitShape;

/* Packed parameter table: 5 + 12*depth device pointers, order (docs/VIT_SPEC.md):
 *   0 patch_embed.kernel bf16 [D, P*P*C]   1 patch_embed.bias f32 [D]   2 pos_embed f32 [T, D]
 *   per block l at 3 + 12*l: ln1.scale, ln1.bias (f32 [D]); qkv.kernel bf16 [3D, D]; qkv.bias f32 [3D];
 *     out.kernel bf16 [D, D]; out.bias f32 [D]; ln2.scale, ln2.bias; fc1.kernel bf16 [mlp, D];
 *     fc1.bias f32 [mlp]; fc2.kernel bf16 [D, mlp]; fc2.bias f32 [D]
 *   last two: encoder_norm.scale, encoder_norm.bias (f32 [D])
 * Matrices are stored [out, in] (the transpose of the flax Dense kernel).  The gradient table of
 * vitmarl_vit_bwd has the same order and shapes, all fp32. */
int vitmarl_vit_num_params(const VitmarlVitShape* s);
long long vitmarl_vit_param_elems(const VitmarlVitShape* s, int index, int* is_bf16_matrix);
/* Bytes of caller-owned activation workspace for fwd (save_for_bwd = 0) or fwd+bwd (1). */
size_t vitmarl_vit_workspace_bytes(const VitmarlVitShape* s, int save_for_bwd);

/* Flag for save_for_bwd: x is already the patch matrix [B*T, P*P*C] bf16 (e.g. rendered by vitmarl_env_step with
 * VITMARL_IMG_BF16_PATCHES_OF(P)), so no patchify pass runs.  With save_for_bwd = 1 the matrix is copied into the workspace
 * (the backward pass reads it for the patch-embedding weight gradient); vitmarl_vit_bwd's dx stays d(image) [B,H,W,C]. */
#define VITMARL_VIT_INPUT_PATCHES 4

/* y[B,D] (fp32) = ViT(x[B,H,W,C] bf16).  Replaces `module.apply({'params': p}, x)`.
 * save_for_bwd: 0 inference, 1 keep the activations vitmarl_vit_bwd needs, 2 inference that REUSES the folded
 * parameters a previous call (mode 0, same workspace, same batch size, unchanged `params` contents) left in the
 * workspace -- the rollout loop between two optimiser updates; skips the two parameter-fold launches. */
int vitmarl_vit_fwd(void* stream, const VitmarlVitShape* s, const void* const* params,
                    const void* x, float* y, void* workspace, size_t workspace_bytes, int save_for_bwd);

/* VJP: gradients of <y, dy> w.r.t. every packed parameter (fp32, zeroed here) and, if dx != NULL,
 * w.r.t. the input (bf16 [B,H,W,C]).  `workspace` is the one filled by vitmarl_vit_fwd(..., 1). */
int vitmarl_vit_bwd(void* stream, const VitmarlVitShape* s, const void* const* params,
                    void* workspace, size_t workspace_bytes, const float* dy,
                    void* const* dparams, void* dx);

/* Per-call options of vitmarl_vit_fwd_ex / vitmarl_vit_bwd_ex.  The library keeps NO process-global mutable state: every
 * switch travels with the call, so XLA may invoke the handlers concurrently from several threads / devices.  Zero-initialise,
 * then set fields; integer switches use -1 (or 0 pointers) for "default".  NULL options = all defaults. */
typedef struct VitmarlVitOptions {
  int fused;                  /* -1 default (1): fused per-block tcgen05 kernels where the shape allows; 0: the unfused kernel sequence */
  int gemm_2cta;              /* -1 default (1): CTA-pair GEMMs; 0: 1-CTA kernels only                                      */
  int pdl;                    /* -1 default (1): programmatic dependent launch of the fused block kernels                   */
  int attn_flags;             /* -1 default: tuning switches of the fused attention block (FusedAttn2Params::flags)         */
  void* timing;               /* handle from vitmarl_timing_create: every launch of this call is bracketed by CUDA events   */
  long long* debug_timeline;  /* device buffer (>= 512 int64) receiving clock64() phase stamps of the fused block kernels    */
  /* backward only */
  void* grads_flat;           /* if the gradient table is carved out of ONE buffer: its base ...                            */
  size_t grads_flat_bytes;    /* ... and size, zeroed with a single memset instead of one per tensor                        */
  int accumulate;             /* 1: do NOT zero the gradient table first -- this call's gradients are added to its contents
                                 (micro-batches of one minibatch accumulate in place; every dW / bias / LayerNorm gradient is
                                 produced by fp32 reductions anyway).  0 / -1: zero first (default)                         */
  void* const* bucket_events; /* cudaEvent_t[vitmarl_vit_num_buckets()] or NULL.  Event b is recorded on `stream` as soon as
                                 gradient bucket b is final: 0 = encoder_norm, 1 + (depth-1-l) = block l (its 12 tensors),
                                 depth+1 = patch_embed + pos_embed.  The caller starts each bucket's all-reduce on a side stream
                                 behind its event, overlapping the rest of the backward pass (the pmean of
                                 jaxrl/MARL/ippo_rnn_JAXMARL_pmap.py:564-565).                                              */
} VitmarlVitOptions;

int vitmarl_vit_fwd_ex(void* stream, const VitmarlVitShape* s, const void* const* params,
                       const void* x, float* y, void* workspace, size_t workspace_bytes, int save_for_bwd,
                       const VitmarlVitOptions* options);
int vitmarl_vit_bwd_ex(void* stream, const VitmarlVitShape* s, const void* const* params,
                       void* workspace, size_t workspace_bytes, const float* dy,
                       void* const* dparams, void* dx, const VitmarlVitOptions* options);
/* Number of gradient buckets (depth + 2) of vitmarl_vit_bwd_ex's bucket_events. */
int vitmarl_vit_num_buckets(const VitmarlVitShape* s);

/* Measurement: a caller-owned log of CUDA-event timings (on the launch stream) of every kernel launch issued by the calls that
 * carry the handle in VitmarlVitOptions::timing.  read() synchronises on the last logged launch and returns per-category
 * totals -- arrays of 8: 0 gemm, 1 fused MLP block, 2 fused attention block, 3 attention, 4 layernorm, 5 other (patchify,
 * pool, parameter folding), 6 dW GEMMs, 7 dX GEMMs -- and the algorithmic FLOPs of the tensor-core launches. */
void* vitmarl_timing_create(void);
void vitmarl_timing_destroy(void* timing);
int vitmarl_timing_reset(void* timing);
int vitmarl_timing_read(void* timing, double* ms8, long long* n8, double* flops);

/* ---- policy head behind the encoder ------------------------------------------------ */

/* The Dense / GRU pieces of ActorCriticRNN.__call__(hidden, (obs, dones)) (jaxrl/MARL/ippo_rnn_JAXMARL.py:48-115), fp32:
 *   y[R,N] = act( [x0 | x1] . W + bias ),  x0 [R,K0] (row pitch ldx0), x1 [R,K1] or NULL (K1 = 0), W [K0+K1, N] in the flax Dense
 *   kernel layout [in, out], bias [N] or NULL.  The two-source form is the first Dense on concat(vector observation, ViT encoding). */
#define VITMARL_ACT_NONE 0
#define VITMARL_ACT_RELU 1
int vitmarl_dense_f32(void* stream, int R, int K0, int K1, int N, const float* x0, int ldx0, const float* x1, int ldx1,
                      const float* W, const float* bias, int act, float* y, int ldy);
/* flax.linen.GRUCell step with the ScannedRNN reset (ippo_rnn_JAXMARL.py:48-66): rows with reset[r] != 0 start from a zero carry.
 *   gi [R,3H] = x . [W_ir | W_iz | W_in] + [b_ir | b_iz | b_in],  gh [R,3H] = h . [W_hr | W_hz | W_hn] (no bias),  b_hn [H],
 *   r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r * (gh_n + b_hn)), h_out = (1 - z) n + z h.
 * reset: uint8 [R] or NULL; h_out may alias h. */
int vitmarl_gru_cell_f32(void* stream, int R, int H, const float* gi, const float* gh, const float* b_hn, const float* h,
                         const uint8_t* reset, float* h_out);

/* `_calculate_gae` of the PPO trainers (ippo_rnn_JAXMARL.py:372-394), the reverse scan over the trajectory, for all B = NUM_ENVS *
 * agents columns at once: reward / value / done [S,B] (time-major, done uint8 = transition.global_done), last_val [B] ->
 * advantages [S,B], targets = advantages + value [S,B] (nullable).  fp32 with the reference's order of operations, no FMA
 * contraction: bit-identical to the NumPy restatement of the scan. */
int vitmarl_gae_f32(void* stream, int S, int B, float gamma, float gae_lambda, const float* reward, const float* value,
                    const uint8_t* done, const float* last_val, float* advantages, float* targets);

#ifdef __cplusplus
}
#endif
#endif /* VITMARL_B200_H_ */
