// Callers on either side of the order-book kernel: integer gather / select glue
// of MARLEnv.step_env that touches the books, trades and the shared message array.  HBM-bound byte work: one thread
// per 32-byte message row (2 x 16-byte accesses), grids sized from the SM count.
//   vitmarl_get_cancel_msgs   job.getCancelMsgs                         JaxOrderBookArrays.py:756-782
//   vitmarl_get_agent_trades  job.get_agent_trades                      JaxOrderBookArrays.py:824-831
//   vitmarl_build_step_msgs   BaseLOBEnv._get_data_messages + order-id renumbering + (given) shuffle + concatenate
//                             base_env.py:341-371, marl_env.py:272-344
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "vit_kernels.h"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

// one warp per environment: the first `size` rows (ascending) whose trader id equals agent_id
__global__ void __launch_bounds__(128) cancel_msgs_kernel(int E, int N, int size, const int32_t* __restrict__ book, int agent_id, int side,
                                                           const int32_t* __restrict__ cancel_time, int32_t* __restrict__ out) {
  const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= E) return;
  const int32_t* b = book + (size_t)e * N * 6;
  const int t_s = cancel_time[2 * e], t_ns = cancel_time[2 * e + 1];
  int found = 0;
  for (int base = 0; base < N && found < size; base += 32) {
    const int r = base + lane;
    const bool hit = r < N && b[r * 6 + 3] == agent_id;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    const int pos = found + __popc(m & ((1u << lane) - 1u));
    if (hit && pos < size) {
      int4* o = reinterpret_cast<int4*>(out + ((size_t)e * size + pos) * 8);
      o[0] = make_int4(2, side, b[r * 6 + 1], b[r * 6 + 0]);
      o[1] = make_int4(b[r * 6 + 2], b[r * 6 + 3], t_s, t_ns);
    }
    found += __popc(m);
  }
  // fill_value=-1 indexes the appended all-zero row (JOBA:770-771)
  for (int pos = min(found, size) + lane; pos < size; pos += 32) {
    int4* o = reinterpret_cast<int4*>(out + ((size_t)e * size + pos) * 8);
    o[0] = make_int4(2, side, 0, 0);
    o[1] = make_int4(0, 0, t_s, t_ns);
  }
}

__global__ void agent_trades_kernel(size_t rows, const int4* __restrict__ trades, int agent_id, int4* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
    int4 a = __ldg(trades + 2 * i), b = __ldg(trades + 2 * i + 1);
    const bool executed = a.x >= 0;                            // JOBA:827
    if (!executed) { a = make_int4(0, 0, 0, 0); b = a; }
    const bool mine = (agent_id == b.z) || (agent_id == b.w);   // columns 6, 7 (JOBA:829)
    if (!mine) { a = make_int4(0, 0, 0, 0); b = a; }
    out[2 * i] = a;
    out[2 * i + 1] = b;
  }
}

struct StepMsgParams {
  int E, n_total, n_data, Mc, Ma;
  const int32_t* message_data; const int32_t* start_index; const int32_t* step_counter; const int32_t* end_time_s;
  const int32_t* cancel_msgs; const int32_t* action_msgs; const int32_t* perm; const int32_t* order_id_counter;
  int32_t* combined; int32_t* new_counter;
};

__global__ void step_msgs_kernel(const StepMsgParams p) {
  const int M = p.Mc + p.Ma + p.n_data;
  const size_t total = (size_t)p.E * M;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / M), j = (int)(i % M);
    int4 a, b;
    if (j < p.Mc) {                                            // cancel messages first (marl_env.py:344)
      const int4* s = reinterpret_cast<const int4*>(p.cancel_msgs + ((size_t)e * p.Mc + j) * 8);
      a = __ldg(s); b = __ldg(s + 1);
    } else if (j < p.Mc + p.Ma) {                              // action messages: renumber, then shuffle (marl_env.py:314-324)
      const int k = j - p.Mc;
      const int src = p.perm ? p.perm[(size_t)e * p.Ma + k] : k;
      const int4* s = reinterpret_cast<const int4*>(p.action_msgs + ((size_t)e * p.Ma + src) * 8);
      a = __ldg(s); b = __ldg(s + 1);
      b.x = wsub(p.order_id_counter[e], src);                  // new_order_ids = counter - arange(Ma), applied before the shuffle
    } else {                                                   // data messages (base_env.py:341-371)
      const int k = j - p.Mc - p.Ma;
      long long off = (long long)p.start_index[e] + (long long)p.n_data * p.step_counter[e];
      const long long hi = (long long)p.n_total - p.n_data;    // lax.dynamic_slice clamps the start index
      off = off < 0 ? 0 : (off > hi ? hi : off);
      const int4* s = reinterpret_cast<const int4*>(p.message_data + ((size_t)off + k) * 8);
      a = __ldg(s); b = __ldg(s + 1);
      if (p.end_time_s && b.z >= p.end_time_s[e]) { a = make_int4(0, 0, 0, 0); b.x = 0; b.y = 0; }   // fixed_time: zero all but the time
    }
    int4* d = reinterpret_cast<int4*>(p.combined + i * 8);
    d[0] = a; d[1] = b;
    if (j == 0 && p.new_counter) p.new_counter[e] = wsub(p.order_id_counter[e], p.Ma);
  }
}

// ---- auto-reset (marl_env.py:737-766): where done[e], the world-state leaves of env e are replaced by the reset state of its
// sampled data window (base_env.py:215-231 index_tree(init_states_array, idx); marl_env.py:186-208 best bid / ask tiled over the
// message slots, mid = float32((best_bid + best_ask) / 2), trades of the loaded state).  One warp per environment; environments
// that are not done are untouched (the reference materialises a full reset state for every env and selects).
struct ResetParams {
  int E, N, T, M, n_windows;
  int32_t* bad_window;
  const int32_t* done; const int32_t* window;
  const int32_t* init_asks; const int32_t* init_bids; const int32_t* init_trades; const int32_t* init_best_asks; const int32_t* init_best_bids;
  int32_t* asks; int32_t* bids; int32_t* trades; int32_t* best_asks; int32_t* best_bids; float* mid;
};

__global__ void __launch_bounds__(128) auto_reset_kernel(const ResetParams p) {
  const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= p.E || p.done[e] == 0) return;
  const int wi = p.window[e];
  if (wi < 0 || wi >= p.n_windows) {                                // never index the init tables out of bounds
    if (lane == 0 && p.bad_window) *p.bad_window = 1;
    return;
  }
  const size_t w = (size_t)wi;
  const int side_v = p.N * 6 / 2;                                   // int2 elements per book side (N * 6 is even)
  const int2* sa = reinterpret_cast<const int2*>(p.init_asks + w * p.N * 6);
  const int2* sb = reinterpret_cast<const int2*>(p.init_bids + w * p.N * 6);
  int2* da = reinterpret_cast<int2*>(p.asks + (size_t)e * p.N * 6);
  int2* db = reinterpret_cast<int2*>(p.bids + (size_t)e * p.N * 6);
  for (int i = lane; i < side_v; i += 32) { da[i] = __ldg(sa + i); db[i] = __ldg(sb + i); }
  int4* dt = reinterpret_cast<int4*>(p.trades + (size_t)e * p.T * 8);
  if (p.init_trades) {
    const int4* st = reinterpret_cast<const int4*>(p.init_trades + w * p.T * 8);
    for (int i = lane; i < p.T * 2; i += 32) dt[i] = __ldg(st + i);
  } else {
    for (int i = lane; i < p.T * 2; i += 32) dt[i] = make_int4(-1, -1, -1, -1);
  }
  const int2 ba = __ldg(reinterpret_cast<const int2*>(p.init_best_asks + w * 2));
  const int2 bb = __ldg(reinterpret_cast<const int2*>(p.init_best_bids + w * 2));
  int2* oa = reinterpret_cast<int2*>(p.best_asks + (size_t)e * p.M * 2);
  int2* ob = reinterpret_cast<int2*>(p.best_bids + (size_t)e * p.M * 2);
  for (int i = lane; i < p.M; i += 32) { oa[i] = ba; ob[i] = bb; }
  if (lane == 0 && p.mid) p.mid[e] = __int2float_rn(wadd(bb.x, ba.x)) / 2.0f;   // jnp.float32((best_bid[0] + best_ask[0]) / 2), x64 disabled
}

// ---------------------------------------------------------------- per-step trade reductions of the reward functions
// vision_env.py:2076-2078 (signed executed quantity), :2156-2163 (agentQuant), :2191 (c_rl = sum(price // tick * |qty|)),
// mm_env.py:1906-1936 (_extract_agent_trade_stats: buyQuant / sellQuant / TradedVolume / inventory_delta).
// One warp per environment; int32 wrap-around arithmetic like XLA; out[e] = 8 int32:
//   [sum qty, sum |qty|, c_rl, buyQuant, sellQuant, TradedVolume, inventory_delta, sum |qty| of the OTHER executed trades]
__global__ void __launch_bounds__(128) trade_stats_kernel(int E, int T, const int4* __restrict__ trades, int agent_id, int tick, int4* __restrict__ out) {
  const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= E) return;
  const int4* tr = trades + (size_t)e * T * 2;
  unsigned s_q = 0, s_abs = 0, s_rev = 0, s_buy = 0, s_sell = 0, s_other = 0;
  for (int r = lane; r < T; r += 32) {
    const int4 a = __ldg(tr + 2 * r), b = __ldg(tr + 2 * r + 1);   // [price, qty, pass_oid, agr_oid], [t_s, t_ns, pass_tid, agr_tid]
    const bool executed = a.x >= 0;
    const int price = executed ? a.x : 0, qty = executed ? a.y : 0, ptid = executed ? b.z : 0, atid = executed ? b.w : 0;
    const bool mine = agent_id == ptid || agent_id == atid;        // evaluated on the zeroed row, as the reference does
    const unsigned aq = qty < 0 ? 0u - (unsigned)qty : (unsigned)qty;   // jnp.abs wraps for INT_MIN
    if (mine) {
      s_q += (unsigned)qty;
      s_abs += aq;
      s_rev += (unsigned)(price / tick) * aq;                      // price >= 0 here: floor division == C division
      const bool buy = (qty >= 0 && agent_id == ptid) || (qty < 0 && agent_id == atid);
      const bool sell = (qty < 0 && agent_id == ptid) || (qty >= 0 && agent_id == atid);
      if (buy) s_buy += aq;
      if (sell) s_sell += aq;
    } else {
      s_other += aq;
    }
  }
  s_q = __reduce_add_sync(0xffffffffu, s_q); s_abs = __reduce_add_sync(0xffffffffu, s_abs); s_rev = __reduce_add_sync(0xffffffffu, s_rev);
  s_buy = __reduce_add_sync(0xffffffffu, s_buy); s_sell = __reduce_add_sync(0xffffffffu, s_sell); s_other = __reduce_add_sync(0xffffffffu, s_other);
  if (lane == 0) {
    out[2 * e] = make_int4((int)s_q, (int)s_abs, (int)s_rev, (int)s_buy);
    out[2 * e + 1] = make_int4((int)s_sell, (int)(s_buy + s_sell), (int)(s_buy - s_sell), (int)s_other);
  }
}

// ---------------------------------------------------------------- _filter_messages (vision_env.py:622-684 = mm_env.py:509-571)
// One thread per environment, n <= 32 action rows and as many cancel rows.  Closed forms of the reference's index
// gymnastics: where(mask, size=n, fill=-1) lists the set rows in ascending order; rank_rev(mask)[i] = (#set before i) for
// a set row and (#set + #unset before i) for an unset one; rel[t] = (c[t] >= a[t]) * a[t] pairs the t-th matching action
// with the t-th matching cancellation (0 quantities past the end of either list).
constexpr int kMaxFilterRows = 32;
__global__ void __launch_bounds__(128) filter_msgs_kernel(int E, int n, const int32_t* __restrict__ action_in, const int32_t* __restrict__ cnl_in,
                                                          int32_t* __restrict__ action_out, int32_t* __restrict__ cnl_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int4* ain = reinterpret_cast<const int4*>(action_in) + (size_t)e * n * 2;
  const int4* cin = reinterpret_cast<const int4*>(cnl_in) + (size_t)e * n * 2;
  int pa[kMaxFilterRows], qa[kMaxFilterRows], pc[kMaxFilterRows], qc[kMaxFilterRows];
  for (int i = 0; i < n; ++i) { const int4 a = ain[2 * i], c = cin[2 * i]; qa[i] = a.z; pa[i] = a.w; qc[i] = c.z; pc[i] = c.w; }
  unsigned am = 0, cm = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (pc[j] == pa[i] && pa[i] != 0) { am |= 1u << i; cm |= 1u << j; }
  int a[kMaxFilterRows], c[kMaxFilterRows];       // quantities of the matching rows in order, 0-padded
  int ka = 0, kc = 0;
  for (int i = 0; i < n; ++i) { a[i] = 0; c[i] = 0; }
  for (int i = 0; i < n; ++i) { if (am >> i & 1) a[ka++] = qa[i]; if (cm >> i & 1) c[kc++] = qc[i]; }
  int rel[kMaxFilterRows];
  for (int t = 0; t < n; ++t) rel[t] = c[t] >= a[t] ? a[t] : 0;
  int4* aout = reinterpret_cast<int4*>(action_out) + (size_t)e * n * 2;
  int4* cout_ = reinterpret_cast<int4*>(cnl_out) + (size_t)e * n * 2;
  int sa = 0, ua = 0, sc = 0, uc = 0;              // set / unset rows seen so far
  for (int i = 0; i < n; ++i) {
    int4 r0 = ain[2 * i], r1 = ain[2 * i + 1];
    const bool set = am >> i & 1;
    const int rk = set ? sa : ka + ua;
    if (set) ++sa; else ++ua;
    r0.z = wsub(r0.z, rel[rk]);
    if (r0.z == 0) { r0 = make_int4(0, 0, 0, 0); r1 = r0; }   // actions netted to zero become dummy messages
    aout[2 * i] = r0; aout[2 * i + 1] = r1;
  }
  for (int j = 0; j < n; ++j) {
    int4 r0 = cin[2 * j];
    const int4 r1 = cin[2 * j + 1];
    const bool set = cm >> j & 1;
    const int rk = set ? sc : kc + uc;
    if (set) ++sc; else ++uc;
    r0.z = wsub(r0.z, rel[rk]);
    cout_[2 * j] = r0; cout_[2 * j + 1] = r1;
  }
}

// ---- default action -> message tables of the two agents -----------------------------------------------------------
// ExecutionAgent._getActionMsgs_fixedQuant_complex (vision_env.py:1046-1142) and MarketMakingAgent._getActionMsgs_spread_skew
// (mm_env.py:1352-1491): one thread per environment.  Integer parts are exact; the float parts reproduce JAX with x64 disabled --
// int32 operands converted to float32, ONE rounding per operation (no FMA contraction), `//` on floats as jax.numpy.floor_divide's
// _float_divmod (fmod, subtract, divide, sign fix-up, round half away from zero), float -> int32 by truncation.
__device__ __forceinline__ float jnp_floor_divide_f32(float x, float y) {
  const float mod = fmodf(x, y);
  float div = __fdiv_rn(__fsub_rn(x, mod), y);
  const bool ind = (mod != 0.0f) && ((y < 0.0f) != (mod < 0.0f));     // sign(y) != sign(mod) with both non-zero
  if (ind) div = __fsub_rn(div, 1.0f);
  return copysignf(floorf(__fadd_rn(fabsf(div), 0.5f)), div);          // lax.round: half away from zero
}
__device__ __forceinline__ int f32_to_i32_trunc(float x) {              // convert_element_type: toward zero, saturating, NaN -> 0
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (int)0x80000000;
  return (int)x;
}
__device__ __forceinline__ int floordiv_i32(int a, int b) {            // jnp.floor_divide on int32
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

struct ExecActionParams {
  int E; const int32_t* action; const int32_t* best_ask; const int32_t* best_bid; int price_stride;
  const int32_t* is_sell; const int32_t* task_to_execute; const int32_t* quant_executed; const int32_t* time;
  int trader_id, tick, n_ticks, fixed_quant, delay, placeholder; int32_t* out;
};
__global__ void __launch_bounds__(128) exec_action_msgs_kernel(const ExecActionParams p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.E) return;
  const int tick = p.tick;
  const int best_ask = (int)((unsigned)floordiv_i32(p.best_ask[(size_t)e * p.price_stride], tick) * (unsigned)tick);
  const int best_bid = (int)((unsigned)floordiv_i32(p.best_bid[(size_t)e * p.price_stride], tick) * (unsigned)tick);
  const bool sell = p.is_sell[e] != 0;
  const int sum = (int)((unsigned)best_bid + (unsigned)best_ask);
  int lv[4];
  if (!sell) {
    lv[0] = best_ask;
    lv[1] = (int)((unsigned)floordiv_i32(floordiv_i32(sum, 2), tick) * (unsigned)tick);
    lv[2] = best_bid;
    lv[3] = (int)((unsigned)best_bid - (unsigned)(tick * p.n_ticks));
  } else {
    lv[0] = best_bid;
    const float half = __fdiv_rn((float)sum, 2.0f);
    lv[1] = f32_to_i32_trunc(__fmul_rn(ceilf(jnp_floor_divide_f32(half, (float)tick)), (float)tick));
    lv[2] = best_ask;
    lv[3] = (int)((unsigned)best_ask + (unsigned)(tick * p.n_ticks));
  }
  // quant_array (vision_env.py:1098-1112): action 0 = nothing; 1 + 4 k + c = column c with multiplier {1, 2, 5}[k]
  const int a = p.action[e];
  int q[4] = {0, 0, 0, 0};
  if (a >= 1 && a <= 12) { const int k = (a - 1) >> 2, c = (a - 1) & 3; q[c] = (int)((unsigned)(k == 0 ? 1 : (k == 1 ? 2 : 5)) * (unsigned)p.fixed_quant); }
  const int quant_left = (int)((unsigned)p.task_to_execute[e] - (unsigned)p.quant_executed[e]);
  const int total = (int)((unsigned)q[0] + (unsigned)q[1] + (unsigned)q[2] + (unsigned)q[3]);
  if (total <= quant_left) {
#pragma unroll
    for (int c = 0; c < 4; ++c) q[c] = f32_to_i32_trunc((float)q[c]);            // both branches of the jnp.where pass through float32
  } else {
    q[0] = f32_to_i32_trunc(floorf((float)quant_left)); q[1] = q[2] = q[3] = 0;
  }
  const int side = sell ? -1 : 1;
  const int ts = (int)((unsigned)p.time[2 * e] + (unsigned)p.delay), tn = (int)((unsigned)p.time[2 * e + 1] + (unsigned)p.delay);
  int4* o = reinterpret_cast<int4*>(p.out + (size_t)e * 32);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    o[2 * c] = make_int4(1, side, q[c], lv[c]);
    o[2 * c + 1] = make_int4(p.placeholder, p.trader_id, ts, tn);
  }
}

struct MmActionParams {
  int E; const int32_t* action; const int32_t* best_ask; const int32_t* best_bid; int price_stride; const int32_t* time;
  int trader_id, tick; float spread_mult, skew_mult; int spread_type_mult, fixed_quant, delay, placeholder; int32_t* out;
};
__global__ void __launch_bounds__(128) mm_action_msgs_kernel(const MmActionParams p) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.E) return;
  const int tick = p.tick;
  const float ftick = (float)tick;
  const int best_ask = (int)((unsigned)floordiv_i32(p.best_ask[(size_t)e * p.price_stride], tick) * (unsigned)tick);
  const int best_bid = (int)((unsigned)floordiv_i32(p.best_bid[(size_t)e * p.price_stride], tick) * (unsigned)tick);
  const float mid = __fdiv_rn((float)(int)((unsigned)best_ask + (unsigned)best_bid), 2.0f);
  const int spread = (int)((unsigned)best_ask - (unsigned)best_bid);
  const int a = p.action[e];
  const int spread_type = floordiv_i32(a, 3), skew_type = a - 3 * spread_type;      // jnp // and % (floor semantics)
  const float new_spread = __fmul_rn((float)spread, spread_type == 0 ? 1.0f : p.spread_mult);
  const float skew = skew_type == 0 ? -p.skew_mult : (skew_type == 1 ? 0.0f : p.skew_mult);
  const float skewed_mid = __fadd_rn(mid, __fmul_rn(skew, p.spread_type_mult ? new_spread : ftick));
  const float half = jnp_floor_divide_f32(new_spread, 2.0f);
  const float bid = __fmul_rn(jnp_floor_divide_f32(__fsub_rn(skewed_mid, half), ftick), ftick);
  const float ask = __fmul_rn(jnp_floor_divide_f32(__fadd_rn(skewed_mid, half), ftick), ftick);
  const int ts = (int)((unsigned)p.time[2 * e] + (unsigned)p.delay), tn = (int)((unsigned)p.time[2 * e + 1] + (unsigned)p.delay);
  int4* o = reinterpret_cast<int4*>(p.out + (size_t)e * 16);
  o[0] = make_int4(1, 1, p.fixed_quant, f32_to_i32_trunc(bid));
  o[1] = make_int4(p.placeholder, p.trader_id, ts, tn);
  o[2] = make_int4(1, -1, p.fixed_quant, f32_to_i32_trunc(ask));
  o[3] = make_int4(p.placeholder, p.trader_id, ts, tn);
}

}  // namespace vitmarl

using namespace vitmarl;

extern "C" int vitmarl_get_cancel_msgs(void* stream, int E, int N, int size, const int32_t* bookside, int agent_id, int side,
                                       const int32_t* cancel_time, int32_t* out) {
  if (E == 0 || size == 0) return VITMARL_OK;
  if (E < 0 || N < 1 || size < 0 || !bookside || !cancel_time || !out || (reinterpret_cast<uintptr_t>(out) & 15)) return VITMARL_EINVAL;
  cancel_msgs_kernel<<<(E + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(E, N, size, bookside, agent_id, side, cancel_time, out);
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_get_agent_trades(void* stream, int E, int T, const int32_t* trades, int agent_id, int32_t* out) {
  const size_t rows = (size_t)E * T;
  if (rows == 0) return VITMARL_OK;
  if (E < 0 || T < 0 || !trades || !out || ((reinterpret_cast<uintptr_t>(trades) | reinterpret_cast<uintptr_t>(out)) & 15)) return VITMARL_EINVAL;
  const int grid = (int)((rows + 255) / 256 < (size_t)num_sms() * 8 ? (rows + 255) / 256 : (size_t)num_sms() * 8);
  agent_trades_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, reinterpret_cast<const int4*>(trades), agent_id,
                                                                           reinterpret_cast<int4*>(out));
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_filter_messages(void* stream, int E, int n, const int32_t* action_msgs, const int32_t* cnl_msgs,
                                       int32_t* action_out, int32_t* cnl_out) {
  if (E == 0 || n == 0) return VITMARL_OK;
  if (E < 0 || n < 0 || n > kMaxFilterRows || !action_msgs || !cnl_msgs || !action_out || !cnl_out) return VITMARL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(action_msgs) | reinterpret_cast<uintptr_t>(cnl_msgs) | reinterpret_cast<uintptr_t>(action_out) |
       reinterpret_cast<uintptr_t>(cnl_out)) & 15) return VITMARL_EINVAL;
  filter_msgs_kernel<<<(E + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(E, n, action_msgs, cnl_msgs, action_out, cnl_out);
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_agent_trade_stats(void* stream, int E, int T, const int32_t* trades, int agent_id, int tick_size, int32_t* out) {
  if (E == 0) return VITMARL_OK;
  if (E < 0 || T < 0 || tick_size < 1 || !trades || !out || ((reinterpret_cast<uintptr_t>(trades) | reinterpret_cast<uintptr_t>(out)) & 15)) return VITMARL_EINVAL;
  trade_stats_kernel<<<(E + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(E, T, reinterpret_cast<const int4*>(trades), agent_id, tick_size,
                                                                               reinterpret_cast<int4*>(out));
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_build_step_msgs(void* stream, int E, int n_total, int n_data, int Mc, int Ma, const int32_t* message_data,
                                       const int32_t* start_index, const int32_t* step_counter, const int32_t* end_time_s,
                                       const int32_t* cancel_msgs, const int32_t* action_msgs, const int32_t* perm,
                                       const int32_t* order_id_counter, int32_t* combined, int32_t* new_order_id_counter) {
  if (E == 0) return VITMARL_OK;
  if (E < 0 || n_data < 0 || Mc < 0 || Ma < 0 || n_total < n_data || !combined) return VITMARL_EINVAL;
  if ((n_data > 0 && (!message_data || !start_index || !step_counter)) || (Mc > 0 && !cancel_msgs) || (Ma > 0 && (!action_msgs || !order_id_counter)))
    return VITMARL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(combined) | reinterpret_cast<uintptr_t>(message_data) | reinterpret_cast<uintptr_t>(cancel_msgs) |
       reinterpret_cast<uintptr_t>(action_msgs)) & 15) return VITMARL_EINVAL;
  if (Mc + Ma + n_data == 0) return VITMARL_OK;
  StepMsgParams p{E, n_total, n_data, Mc, Ma, message_data, start_index, step_counter, end_time_s, cancel_msgs, action_msgs, perm,
                  order_id_counter, combined, new_order_id_counter};
  const size_t total = (size_t)E * (Mc + Ma + n_data);
  const int grid = (int)((total + 255) / 256 < (size_t)num_sms() * 8 ? (total + 255) / 256 : (size_t)num_sms() * 8);
  step_msgs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_auto_reset(void* stream, int E, int N, int T, int M, int n_windows, const int32_t* done, const int32_t* window_index,
                                  const int32_t* init_asks, const int32_t* init_bids, const int32_t* init_trades,
                                  const int32_t* init_best_asks, const int32_t* init_best_bids, int32_t* asks, int32_t* bids,
                                  int32_t* trades, int32_t* best_asks, int32_t* best_bids, float* mid_price, int32_t* bad_window) {
  if (E == 0) return VITMARL_OK;
  if (E < 0 || N < 1 || T < 0 || M < 0 || n_windows < 1 || !done || !window_index || !init_asks || !init_bids || !init_best_asks ||
      !init_best_bids || !asks || !bids || !trades || !best_asks || !best_bids)
    return VITMARL_EINVAL;
  if ((N * 6) % 2) return VITMARL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(trades) | reinterpret_cast<uintptr_t>(init_trades)) & 15) return VITMARL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(asks) | reinterpret_cast<uintptr_t>(bids) | reinterpret_cast<uintptr_t>(init_asks) |
       reinterpret_cast<uintptr_t>(init_bids) | reinterpret_cast<uintptr_t>(best_asks) | reinterpret_cast<uintptr_t>(best_bids) |
       reinterpret_cast<uintptr_t>(init_best_asks) | reinterpret_cast<uintptr_t>(init_best_bids)) & 7)
    return VITMARL_EINVAL;
  ResetParams p{E, N, T, M, n_windows, bad_window, done, window_index, init_asks, init_bids, init_trades, init_best_asks, init_best_bids,
                asks, bids, trades, best_asks, best_bids, mid_price};
  auto_reset_kernel<<<(E + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_exec_action_msgs_fixed_quants_complex(void* stream, int E, const int32_t* action, const int32_t* best_ask_price,
                                                             const int32_t* best_bid_price, int price_stride, const int32_t* is_sell_task,
                                                             const int32_t* task_to_execute, const int32_t* quant_executed,
                                                             const int32_t* time, int trader_id, int tick_size, int n_ticks_in_book,
                                                             int fixed_quant_value, int time_delay_obs_act, int placeholder_order_id,
                                                             int32_t* out) {
  if (E == 0) return VITMARL_OK;
  if (E < 0 || tick_size < 1 || price_stride < 1 || !action || !best_ask_price || !best_bid_price || !is_sell_task || !task_to_execute ||
      !quant_executed || !time || !out || (reinterpret_cast<uintptr_t>(out) & 15))
    return VITMARL_EINVAL;
  ExecActionParams p{E, action, best_ask_price, best_bid_price, price_stride, is_sell_task, task_to_execute, quant_executed, time,
                     trader_id, tick_size, n_ticks_in_book, fixed_quant_value, time_delay_obs_act, placeholder_order_id, out};
  exec_action_msgs_kernel<<<(E + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_cuda(cudaGetLastError());
}

extern "C" int vitmarl_mm_action_msgs_spread_skew(void* stream, int E, const int32_t* action, const int32_t* best_ask_price,
                                                  const int32_t* best_bid_price, int price_stride, const int32_t* time, int trader_id,
                                                  int tick_size, float spread_multiplier, float skew_multiplier, int multiplier_type,
                                                  int fixed_quant_value, int time_delay_obs_act, int placeholder_order_id, int32_t* out) {
  if (E == 0) return VITMARL_OK;
  if (E < 0 || tick_size < 1 || price_stride < 1 || (multiplier_type != 0 && multiplier_type != 1) || !action || !best_ask_price ||
      !best_bid_price || !time || !out || (reinterpret_cast<uintptr_t>(out) & 15))
    return VITMARL_EINVAL;
  MmActionParams p{E, action, best_ask_price, best_bid_price, price_stride, time, trader_id, tick_size, spread_multiplier, skew_multiplier,
                   multiplier_type, fixed_quant_value, time_delay_obs_act, placeholder_order_id, out};
  mm_action_msgs_kernel<<<(E + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_cuda(cudaGetLastError());
}
