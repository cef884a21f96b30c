// Order-book step + observation render for sm_100a: ONE WARP PER ENVIRONMENT.
//
// What it replaces (reference, pure JAX): gymnax_exchange/jaxob/JaxOrderBookArrays.py
//   scan_through_entire_array_save_bidask :720-752  (lax.scan of cond_type_side_save_bidask :617-661)
//   add_order :62-83, cancel_order :93-138, match_order/_match_against_* :171-330,
//   get_best_bid_and_ask_inclQuants :881-898, get_L2_state :1075-1106, get_vision_L2_state :1108-1140
// and gymnax_exchange/jaxen/vision_env.py:2804-2854 (normalize_vision_obs),
//     gymnax_exchange/jaxen/marl_env.py:392-393,466-467,685-711 (_ffill_best_prices, mid price).
//
// B200 design (not a translation of the XLA program, which runs ~40 small fused loops over
// [100,6] arrays per message and, under vmap, executes all five switch branches):
//   * a warp owns one environment for the whole step; book sides / trades are moved with the TMA
//     engine as 1-D bulk copies (cp.async.bulk global<->shared + mbarrier), one instruction per
//     2.4 KB side, and the AoS slab in shared memory stays the master copy of the side: a message
//     touches one row (a few STS by the lane that owns it) and nothing is transposed on the way out;
//   * price / quantity / order id of every row are cached in REGISTERS (row r = j*32 + lane,
//     RPL = ceil(N/32) rows per lane) because every search and reduction runs over them;
//   * messages are read as coalesced 32-byte rows (one message per lane, 2x LDG.128) and
//     broadcast with SHFL; per-message best bid/ask are kept in lane registers and written
//     as coalesced 8-byte rows per 32 messages;
//   * "first index" / min / max / sum over the N rows are REDUX.{MIN,MAX,SUM} + VOTE.BALLOT;
//   * only the branch a message selects is executed (warp-uniform control flow).
// Bit-exactness: every quirk of the reference is reproduced; the
// shortcuts taken on the fast path (touching only the modified row when wiping, checking
// only `price` when looking for an empty row) are guarded by per-side invariants that are
// established at load time and fall back to the literal full-array form otherwise.
//
// HBM bytes per env-step (N=T=100): 2*(2*N*24) + T*32 + M*32 + 2*M*8 = 12800 + 48*M.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"
#include "../../include/vitmarl_b200.h"

namespace vitmarl {

constexpr int kWarpsPerCta = 4;
constexpr int MAXINT = 2147483647;
constexpr unsigned FULL = 0xffffffffu;

struct LobParams {
  int E, N, T, M, n_keep;
  int32_t init_id;
  const int32_t* asks_in; const int32_t* bids_in; const int32_t* trades_in; const int32_t* msgs;
  int32_t* asks_out; int32_t* bids_out; int32_t* trades_out;
  int32_t* best_asks; int32_t* best_bids;
  // fused env-step extras
  int do_step;            // 0: render-only kernel (books are read, not written)
  int do_ffill;           // ffill best prices + mid price
  const int32_t* last_ask_price; const int32_t* last_bid_price;
  int last_stride;        // element stride of the two arrays above (1 = dense [E]; 2*M = column [:, -1, 0] of a best-price track)
  float* mid_out;
  // fused reward reductions over the step's trade log: stats [E, n_stat, 8] for up to 4 trader ids; with them
  // the caller may pass trades_out = NULL and the [T,8] log never leaves the chip
  int n_stat; int32_t stat_ids[4]; int32_t* stats; int stat_tick;
  // render
  int do_render; int n_levels; int tick;
  const float* mid_in;    // render-only: mid price input (nullable -> no norm)
  int32_t* raw; int32_t* l2; float* norm;
  void* image; int img_dtype; int H; int W;
  int bulk_ok;            // 1: TMA bulk copies usable (16-byte aligned, N even)
  const int32_t* time_in; int32_t* time_out; float* delta_time;   // world clock (marl_env.py:406,468,482); all or none
};

// ---------------------------------------------------------------- one book side: registers + shared memory
// The AoS slab [N][6] the TMA engine loaded stays the MASTER copy of the side for the whole step (it is also what
// the TMA store sends back, so nothing is transposed on the way out); price / quantity / order id -- the fields the
// searches and reductions run over -- are cached in registers (row r = j*32 + lane).  Row r of the slab is only ever
// touched by lane r & 31, so the two copies stay coherent without any intra-warp synchronisation.
//
// TWO implementations of the message semantics live in this file:
//   * the FAST path (register cache, templates) assumes both sides are TIDY: every row is either all -1, or has no
//     -1 field, qty > 0 and price / time_s / time_ns != MAXINT.  Then an empty row <=> price == -1, only the touched
//     row can need wiping, the row the price-time priority search returns sits at the cached best price, and the
//     cached best price / volume can be updated incrementally.  Normal LOBSTER-style data never leaves this state.
//   * the LITERAL path (slow_* below) restates JOBA line by line on the shared-memory slabs.  A warp switches to it,
//     for the rest of its step, the moment a side is not tidy at load time or a message could break tidiness.
// Keeping the literal forms out of the hot loop matters: the loop is instruction-FETCH bound (20 warps per SM at
// different program counters; 32 KB L1.5 instruction cache), and the literal branches were a third of its code.
template <int RPL>
struct Side {
  int p[RPL], q[RPL], oid[RPL];
  int32_t* sm;   // [N][6] master copy in shared memory
  int best;      // bids: max(price) | asks: min(price, -1 -> MAXINT)   (over existing rows)
  int vol;       // volume at the reported best price (Q10 semantics)
};

// rows r = j*32 + lane exist for r < N; RPL = ceil(N/32), so only the last j can hold phantom rows
#define VM_EX(j) ((j) < RPL - 1 || (j) * 32 + lane < N)

__device__ __forceinline__ void sm_store_row(int32_t* sm, int r, int a, int b, int c, int d, int e, int f) {
  int2* row = reinterpret_cast<int2*>(sm + r * 6);
  row[0] = make_int2(a, b); row[1] = make_int2(c, d); row[2] = make_int2(e, f);
}

// first row index (ascending) whose predicate holds, -1 if none: one REDUX.MIN over per-lane candidates
template <int RPL>
__device__ __forceinline__ int first_index(const bool (&pred)[RPL], int lane) {
  int cand = MAXINT;
#pragma unroll
  for (int j = RPL - 1; j >= 0; --j)
    if (pred[j]) cand = j * 32 + lane;
  cand = __reduce_min_sync(FULL, cand);
  return cand == MAXINT ? -1 : cand;
}

// JOBA:846-865 + :833-844 -> cached (best, vol) of one side
template <int RPL, bool IS_BID>
__device__ __forceinline__ void refresh_best(Side<RPL>& s, int N, int lane) {
  int m = IS_BID ? (int)0x80000000 : MAXINT;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    if (IS_BID) { if (VM_EX(j)) m = max(m, s.p[j]); }
    else        { m = min(m, s.p[j] == -1 ? MAXINT : s.p[j]); }   // phantom rows hold -1 -> MAXINT: neutral
  }
  m = IS_BID ? __reduce_max_sync(FULL, m) : __reduce_min_sync(FULL, m);
  s.best = m;
  int bp = (!IS_BID && m == MAXINT) ? -1 : m;
  int v = 0;
#pragma unroll
  for (int j = 0; j < RPL; ++j)
    if (VM_EX(j) && s.p[j] == bp) v = wadd(v, s.q[j]);
  s.vol = __reduce_add_sync(FULL, v);
}

template <int RPL, bool IS_BID>
__device__ __forceinline__ int best_price_out(const Side<RPL>& s) {
  return (!IS_BID && s.best == MAXINT) ? -1 : s.best;
}

// JOBA:240-267 on a tidy side whose best price exists (best != -1 / MAXINT): among the rows at the best price the
// earliest (time_s, time_ns), ties -> lowest index.  Times come from the shared-memory copy, only for candidate rows.
template <int RPL>
__device__ __forceinline__ int top_index(const Side<RPL>& s, int lane) {
  const int best = s.best;
  int t[RPL], u[RPL];
  int ms = MAXINT;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    t[j] = MAXINT; u[j] = MAXINT;
    if (s.p[j] == best) {    // phantom rows hold -1 and never match
      const int2 tt = *reinterpret_cast<const int2*>(s.sm + (j * 32 + lane) * 6 + 4);
      t[j] = tt.x; u[j] = tt.y;
    }
    ms = min(ms, t[j]);
  }
  ms = __reduce_min_sync(FULL, ms);
  int mn = MAXINT;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    if (t[j] != ms) u[j] = MAXINT;
    mn = min(mn, u[j]);
  }
  mn = __reduce_min_sync(FULL, mn);   // < MAXINT: tidy rows have time_ns != MAXINT
  bool pred[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) pred[j] = u[j] == mn;
  return first_index<RPL>(pred, lane);
}

// the fields of one message that the book operations use (warp-uniform)
struct Msg { int side, qty, price, oid, tid, ts, tns; };

// JOBA:171-330 : match the incoming order against `book` (Q6, Q7, Q11).  Fresh trade log: rows [0,tr_next) are
// taken, rows [tr_next,T) have column 4 == -1, so the first free slot (Q6) is a counter.
template <int RPL, bool IS_BID_BOOK>
__device__ __forceinline__ int match_against(Side<RPL>& book, const Msg& m, int32_t* tr, int& tr_next, int N, int T, int lane) {
  int qtm = m.qty;
  for (;;) {
    // the row the priority search returns sits AT the cached best price: the loop condition is known before searching
    const int tp = book.best;
    const bool cross = IS_BID_BOOK ? (tp >= m.price) : (tp <= m.price);   // empty ask side: MAXINT <= price never holds
    if (!(cross && qtm > 0 && tp != -1)) break;
    const int top = top_index<RPL>(book, lane);
    const int src = top & 31;
    int2 pq = make_int2(0, 0), ot = make_int2(0, 0);
    if (lane == src) {
      const int2* row = reinterpret_cast<const int2*>(book.sm + top * 6);
      pq = row[0]; ot = row[1];
    }
    const int q_top = __shfl_sync(FULL, pq.y, src);
    const int d = wsub(q_top, qtm);
    const int newq = d > 0 ? d : 0;
    const int filled = wsub(q_top, newq);
    const int e = tr_next < T ? tr_next : T - 1;
    if (lane == src) {
      int4* dst = reinterpret_cast<int4*>(tr + e * 8);
      dst[0] = make_int4(pq.x, wmul(wsub(0, m.side), filled), ot.x, m.oid);
      dst[1] = make_int4(m.ts, m.tns, ot.y, m.tid);
      if (newq <= 0) sm_store_row(book.sm, top, -1, -1, -1, -1, -1, -1);
      else book.sm[top * 6 + 1] = newq;
    }
    if (tr_next < T) tr_next += 1;       // m.ts != -1 on this path, so the slot is taken for good
#pragma unroll
    for (int j = 0; j < RPL; ++j)
      if (j * 32 + lane == top) {
        book.q[j] = newq;
        if (newq <= 0) { book.p[j] = -1; book.q[j] = -1; book.oid[j] = -1; }
      }
    qtm = wsub(qtm, q_top);
    if (newq > 0) book.vol = wsub(book.vol, filled);   // the row stays at the best price
    else refresh_best<RPL, IS_BID_BOOK>(book, N, lane);
  }
  return qtm;
}

// JOBA:62-83 (Q2, Q3) on a tidy side: empty row <=> price == -1.  Returns true when (best, vol) must be recomputed.
template <int RPL, bool IS_BID>
__device__ __forceinline__ bool add_order(Side<RPL>& s, const Msg& m, int qrem, int N, int lane) {
  const int qn = qrem > 0 ? qrem : 0;
  bool pred[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) pred[j] = s.p[j] == -1;      // phantom rows r >= N also hold -1: caught by idx >= N
  int idx = first_index<RPL>(pred, lane);
  const bool found = idx >= 0 && idx < N;
  if (!found) idx = N - 1;                                   // Q2: full side -> last row overwritten
  const bool alive = qn > 0;                                 // write + wipe merged: a zero-quantity row ends up all -1
  const int wp = alive ? m.price : -1, wq = alive ? qn : -1, wo = alive ? m.oid : -1;
  if (lane == (idx & 31))
    sm_store_row(s.sm, idx, wp, wq, wo, alive ? m.tid : -1, alive ? m.ts : -1, alive ? m.tns : -1);
#pragma unroll
  for (int j = 0; j < RPL; ++j)
    if (j * 32 + lane == idx) { s.p[j] = wp; s.q[j] = wq; s.oid[j] = wo; }
  // cached best: the overwritten row was empty, so nothing changes unless the new order rests at / inside the best
  const bool side_nonempty = IS_BID ? (s.best != -1) : (s.best != MAXINT);
  if (!(found && side_nonempty)) return true;
  if (alive) {
    const bool better = IS_BID ? (m.price > s.best) : (m.price < s.best);
    if (better) { s.best = m.price; s.vol = qn; }
    else if (m.price == s.best) s.vol = wadd(s.vol, qn);
  }
  return false;
}

// JOBA:93-138 (Q4, Q5) on a tidy side with msg.qty >= 0.  Returns true when (best, vol) must be recomputed.
template <int RPL, bool IS_BID>
__device__ __forceinline__ bool cancel_order(Side<RPL>& s, const Msg& m, int init_id, int N, int lane) {
  bool pred[RPL];
#pragma unroll
  for (int j = 0; j < RPL; ++j) pred[j] = s.oid[j] == m.oid;       // phantom rows (oid -1) sit above every real row
  int idx = first_index<RPL>(pred, lane);
  if (idx < 0 || idx >= N) {
#pragma unroll
    for (int j = 0; j < RPL; ++j) pred[j] = (s.p[j] == m.price) && (s.oid[j] <= init_id) && (s.q[j] >= m.qty);
    idx = first_index<RPL>(pred, lane);
  }
  if (idx < 0 || idx >= N) idx = N - 1;                               // Q5
  const int src = idx & 31;
  int rp = 0, nq = 0;                                                  // price of the touched row, its new quantity
  if (lane == src) {
    const int2 pq = *reinterpret_cast<const int2*>(s.sm + idx * 6);
    rp = pq.x;
    nq = wsub(pq.y, m.qty);
    if (nq <= 0) sm_store_row(s.sm, idx, -1, -1, -1, -1, -1, -1);
    else s.sm[idx * 6 + 1] = nq;
  }
  rp = __shfl_sync(FULL, rp, src);
  nq = __shfl_sync(FULL, nq, src);
#pragma unroll
  for (int j = 0; j < RPL; ++j)
    if (j * 32 + lane == idx) {
      s.q[j] = nq;
      if (nq <= 0) { s.p[j] = -1; s.q[j] = -1; s.oid[j] = -1; }
    }
  // the cached best only changes when the touched row sat at the best price (or the side had no best)
  const int bp = best_price_out<RPL, IS_BID>(s);
  if (bp == -1) return true;
  if (rp != bp) return false;
  if (nq > 0) { s.vol = wsub(s.vol, m.qty); return false; }            // partial cancel at the best price
  return true;
}

// JOBA:617-661, per lane on ITS message (the code is then broadcast with the message):
//   bits 0-2: branch; bit 3: the fast path may process it -- a limit order whose row would rest cleanly (no -1 among
//   price/oid/tid/ts/tns, price/ts/tns != MAXINT), a cancel with qty >= 0, or a no-op.
__device__ __forceinline__ int decode_message(const int4& m0, const int4& m1) {
  const int t = m0.x, s = m0.y;
  const int idx = (((s == 1 && t == 1) || (s == -1 && t == 4)) ? 1 : 0) + ((s == -1 && (t == 2 || t == 3)) ? 2 : 0) +
                  ((s == 1 && (t == 2 || t == 3)) ? 3 : 0) + ((s == 0 && t == 0) ? 4 : 0);
  const bool clean = !(m0.w == -1 || m1.x == -1 || m1.y == -1 || m1.z == -1 || m1.w == -1) &&
                     m0.w != MAXINT && m1.z != MAXINT && m1.w != MAXINT;
  const bool fast = idx == 4 || (idx <= 1 ? clean : m0.z >= 0);
  return idx | (fast ? 8 : 0);
}

// one message (held by lane i as m0/m1; idx = its branch) applied to a tidy book
template <int RPL>
__device__ __forceinline__ void process_message(Side<RPL>& asks, Side<RPL>& bids, const int4& m0, const int4& m1, int idx,
                                                int i, int32_t* tr, int& tr_next, int N, int T, int init_id, int lane) {
  if (idx == 4) return;    // doNothing :334-355
  Msg m;
  m.qty = __shfl_sync(FULL, m0.z, i); m.price = __shfl_sync(FULL, m0.w, i); m.oid = __shfl_sync(FULL, m1.x, i);
  bool ra = false, rb = false;   // (best, vol) of asks / bids to be recomputed
  if (idx >= 2) {          // ask_cancel :455-478 / bid_cancel :392-415
    if (idx == 2) ra = cancel_order<RPL, false>(asks, m, init_id, N, lane);
    else          rb = cancel_order<RPL, true>(bids, m, init_id, N, lane);
  } else {
    m.tid = __shfl_sync(FULL, m1.y, i); m.ts = __shfl_sync(FULL, m1.z, i); m.tns = __shfl_sync(FULL, m1.w, i);
    if (idx == 0) {        // ask_lim :417-453 (also every (type, side) outside the table, Q8)
      int q = m.qty;
      if (bids.best >= m.price && q > 0) {   // exact quick reject: otherwise the match loop body never runs
        m.side = __shfl_sync(FULL, m0.y, i);
        q = match_against<RPL, true>(bids, m, tr, tr_next, N, T, lane);
      }
      ra = add_order<RPL, false>(asks, m, q, N, lane);
    } else {               // bid_lim :356-391
      int q = m.qty;
      if (asks.best <= m.price && q > 0) {
        m.side = __shfl_sync(FULL, m0.y, i);
        q = match_against<RPL, false>(asks, m, tr, tr_next, N, T, lane);
      }
      rb = add_order<RPL, true>(bids, m, q, N, lane);
    }
  }
  if (ra) refresh_best<RPL, false>(asks, N, lane);
  if (rb) refresh_best<RPL, true>(bids, N, lane);
}

// ---------------------------------------------------------------- literal path (any book, any message)
// JOBA restated on the shared-memory slabs, warp-cooperative, side chosen at run time; every lane returns the same
// values.  Used when a side is not tidy, a message could break tidiness, or the caller supplies a trade log whose
// free rows are not a suffix.  Cross-lane visibility of the single-lane writes comes from __syncwarp().
template <class F>
__device__ __forceinline__ int slow_first(int n, int lane, F pred) {
  for (int base = 0; base < n; base += 32) {
    const int r = base + lane;
    const unsigned b = __ballot_sync(FULL, r < n && pred(r));
    if (b) return base + __ffs(b) - 1;
  }
  return -1;
}

// JOBA:85-90
__device__ __noinline__ void slow_wipe(int32_t* s, int N, int lane) {
  __syncwarp();
  for (int r = lane; r < N; r += 32)
    if (s[r * 6 + 1] <= 0) sm_store_row(s, r, -1, -1, -1, -1, -1, -1);
  __syncwarp();
}

// JOBA:846-865, :833-844 -> raw best (MAXINT for an empty ask side) and the volume at the reported price
__device__ __noinline__ int2 slow_best(const int32_t* s, int N, bool is_bid, int lane) {
  int m = is_bid ? (int)0x80000000 : MAXINT;
  for (int r = lane; r < N; r += 32) {
    const int p = s[r * 6];
    m = is_bid ? max(m, p) : min(m, p == -1 ? MAXINT : p);
  }
  m = is_bid ? __reduce_max_sync(FULL, m) : __reduce_min_sync(FULL, m);
  const int bp = (!is_bid && m == MAXINT) ? -1 : m;
  int v = 0;
  for (int r = lane; r < N; r += 32)
    if (s[r * 6] == bp) v = wadd(v, s[r * 6 + 1]);
  return make_int2(m, __reduce_add_sync(FULL, v));
}

// JOBA:240-267 (Q9)
__device__ __noinline__ int slow_top(const int32_t* s, int N, int best, int lane) {
  int ms = MAXINT;
  for (int r = lane; r < N; r += 32) ms = min(ms, s[r * 6] == best ? s[r * 6 + 4] : MAXINT);
  ms = __reduce_min_sync(FULL, ms);
  int mn = MAXINT;
  for (int r = lane; r < N; r += 32) {
    const int t = s[r * 6] == best ? s[r * 6 + 4] : MAXINT;
    mn = min(mn, t == ms ? s[r * 6 + 5] : MAXINT);
  }
  mn = __reduce_min_sync(FULL, mn);
  const int idx = slow_first(N, lane, [&](int r) {
    const int t = s[r * 6] == best ? s[r * 6 + 4] : MAXINT;
    return (t == ms ? s[r * 6 + 5] : MAXINT) == mn;
  });
  return idx < 0 ? N - 1 : idx;
}

// JOBA:617-661 for one message given as its eight raw fields
__device__ __noinline__ void slow_message(int32_t* asks, int32_t* bids, int32_t* tr, int N, int T, int init_id,
                                          int4 a, int4 b, int lane) {
  const int t = a.x, sd = a.y, qty = a.z, price = a.w, oid = b.x, tid = b.y, ts = b.z, tns = b.w;
  const int idx = (((sd == 1 && t == 1) || (sd == -1 && t == 4)) ? 1 : 0) + ((sd == -1 && (t == 2 || t == 3)) ? 2 : 0) +
                  ((sd == 1 && (t == 2 || t == 3)) ? 3 : 0) + ((sd == 0 && t == 0) ? 4 : 0);
  if (idx == 4) return;
  const bool own_is_bid = (idx & 1) != 0;
  int32_t* own = own_is_bid ? bids : asks;
  int32_t* opp = own_is_bid ? asks : bids;
  __syncwarp();
  if (idx >= 2) {   // cancel_order :93-138 (Q4, Q5)
    int i = slow_first(N, lane, [&](int r) { return own[r * 6 + 2] == oid; });
    if (i < 0) i = slow_first(N, lane, [&](int r) { return own[r * 6] == price && own[r * 6 + 2] <= init_id && own[r * 6 + 1] >= qty; });
    if (i < 0) i = N - 1;
    if (lane == 0) own[i * 6 + 1] = wsub(own[i * 6 + 1], qty);
    slow_wipe(own, N, lane);
    return;
  }
  // match against the opposite side :171-330 (Q6, Q7, Q11)
  const bool book_is_bid = !own_is_bid;
  int qtm = qty;
  for (;;) {
    const int best = slow_best(opp, N, book_is_bid, lane).x;
    const int top = slow_top(opp, N, best, lane);
    const int tp = opp[top * 6];
    const bool cross = book_is_bid ? (tp >= price) : (tp <= price);
    if (!(cross && qtm > 0 && tp != -1)) break;
    const int q_top = opp[top * 6 + 1], o_top = opp[top * 6 + 2], t_top = opp[top * 6 + 3];
    const int d = wsub(q_top, qtm);
    const int newq = d > 0 ? d : 0;
    int e = slow_first(T, lane, [&](int r) { return tr[r * 8 + 4] == -1; });
    if (e < 0) e = T - 1;
    __syncwarp();
    if (lane == 0) {
      int4* dst = reinterpret_cast<int4*>(tr + e * 8);
      dst[0] = make_int4(tp, wmul(wsub(0, sd), wsub(q_top, newq)), o_top, oid);
      dst[1] = make_int4(ts, tns, t_top, tid);
      opp[top * 6 + 1] = newq;
    }
    slow_wipe(opp, N, lane);
    qtm = wsub(qtm, q_top);
  }
  // add_order :62-83 (Q1, Q2, Q3)
  int i = slow_first(N, lane, [&](int r) {
    const int32_t* w = own + r * 6;
    return w[0] == -1 || w[1] == -1 || w[2] == -1 || w[3] == -1 || w[4] == -1 || w[5] == -1;
  });
  if (i < 0) i = N - 1;
  __syncwarp();
  if (lane == 0) sm_store_row(own, i, price, qtm > 0 ? qtm : 0, oid, tid, ts, tns);
  slow_wipe(own, N, lane);
}

// ---------------------------------------------------------------- staging -> registers
template <int RPL>
__device__ __forceinline__ void regs_from_smem(Side<RPL>& s, int32_t* sm, int N, int lane) {
  s.sm = sm;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    s.p[j] = -1; s.q[j] = -1; s.oid[j] = -1;
    if (VM_EX(j)) {
      const int32_t* row = sm + (j * 32 + lane) * 6;
      const int2 a = *reinterpret_cast<const int2*>(row);
      s.p[j] = a.x; s.q[j] = a.y; s.oid[j] = row[2];
    }
  }
}

// Are the `rows` rows of a [rows][6] slab (both sides are contiguous) tidy?  A rolled loop: this runs once per
// environment and its code size counts against the instruction cache the message loop lives in.
__device__ __forceinline__ bool slab_is_tidy(const int32_t* sm, int rows, int lane) {
  bool tidy = true;
#pragma unroll 1
  for (int r = lane; r < rows; r += 32) {
    const int2* row = reinterpret_cast<const int2*>(sm + r * 6);
    const int2 a = row[0], b = row[1], c = row[2];
    // -1 is the largest unsigned value: min == -1 <=> all fields are -1, max == -1 <=> some field is -1
    const unsigned lo = __vimin3_u32(__vimin3_u32(a.x, a.y, b.x), __vimin3_u32(b.y, c.x, c.y), 0xffffffffu);
    const unsigned hi = __vimax3_u32(__vimax3_u32(a.x, a.y, b.x), __vimax3_u32(b.y, c.x, c.y), 0u);
    const bool all = lo == 0xffffffffu, none = hi != 0xffffffffu;
    if (!(all || (none && a.y > 0 && a.x != MAXINT && c.x != MAXINT && c.y != MAXINT))) tidy = false;
  }
  return __all_sync(FULL, tidy);
}

__device__ __forceinline__ void warp_copy_g2s(int32_t* dst, const int32_t* src, int n, int lane) {
#pragma unroll 1
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
}
__device__ __forceinline__ void warp_copy_s2g(int32_t* dst, const int32_t* src, int n, int lane) {
#pragma unroll 1
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// ---------------------------------------------------------------- deterministic log1p
// Same operation sequence as oracle/lob_oracle.{py,c}: IEEE double mul/add/div only
// (__dmul_rn/__dadd_rn/__ddiv_rn are never contracted into FMAs).
__device__ __forceinline__ float vm_log1p_f32(float x) {
  const double LN2_HI = 0x1.62e42fee00000p-1, LN2_LO = 0x1.a39ef35793c76p-33, SQRT2 = 0x1.6a09e667f3bcdp+0;
  const double C[13] = {1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17,
                        1.0 / 19, 1.0 / 21, 1.0 / 23, 1.0 / 25, 1.0 / 27};
  double xd = (double)x;
  if (xd != xd || xd < -1.0) return __int_as_float(0x7fc00000);
  if (xd == -1.0) return __int_as_float(0xff800000);
  if (xd > 1.7e308) return __int_as_float(0x7f800000);
  double y = __dadd_rn(1.0, xd);
  long long bits = __double_as_longlong(y);
  int k = (int)((bits >> 52) & 0x7FF) - 1023;
  bits = (bits & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll;
  double m = __longlong_as_double(bits);
  if (m > SQRT2) { m = __dmul_rn(m, 0.5); k += 1; }
  double f = __dadd_rn(m, -1.0);
  double s = __ddiv_rn(f, __dadd_rn(2.0, f));
  double z = __dmul_rn(s, s);
  double p = C[12];
#pragma unroll
  for (int i = 11; i >= 0; --i) p = __dadd_rn(__dmul_rn(p, z), C[i]);
  p = __dmul_rn(p, z);
  double s2 = __dmul_rn(2.0, s);
  double logm = __dadd_rn(s2, __dmul_rn(s2, p));
  double kd = (double)k;
  double r = __dadd_rn(__dmul_rn(kd, LN2_HI), __dadd_rn(__dmul_rn(kd, LN2_LO), logm));
  return __double2float_rn(r);
}

__device__ __forceinline__ int bar_length(int vol, int W) {
  unsigned u = (unsigned)(vol > 0 ? vol : 0) + 1u;
  int e = 31 - __clz(u);
  unsigned frac2 = ((u << (31 - e)) >> 29) & 3u;
  int l4 = 4 * e + (int)frac2;
  int len = (l4 * W) >> 6;
  return len < W ? len : W;
}

// ---------------------------------------------------------------- stage 2 on a register-resident book
// JOBA:1108-1140 (+ :1075-1106), vision_env.py:2804-2854, docs/RENDER_SPEC.md
template <int RPL>
__device__ __forceinline__ void render_env(const Side<RPL>& asks, const Side<RPL>& bids, const LobParams& P, int e,
                                           bool have_mid, float mid, int32_t* scratch, int lane) {
  const int N = P.N, n = P.n_levels;
  if (P.raw || P.l2 || P.norm) {
    int lp[2] = {-1, -1}, lv[2] = {0, 0};   // this lane's level: price / volume per channel
    int cumv[2] = {0, 0};
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const Side<RPL>& s = ch == 0 ? asks : bids;
      bool have_prev = false;
      int prev = 0;
      int cum = 0;
      for (int l = 0; l < n; ++l) {
        int cur = MAXINT;
        bool cand = false;
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          int key = ch == 0 ? (s.p[j] == -1 ? MAXINT : s.p[j]) : (int)(0u - (unsigned)s.p[j]);
          bool c = VM_EX(j) && (!have_prev || key > prev);
          cand |= c;
          if (c) cur = min(cur, key);
        }
        bool found = __any_sync(FULL, cand);
        cur = __reduce_min_sync(FULL, cur);
        int price;
        if (found) { prev = cur; have_prev = true; price = ch == 0 ? (cur == MAXINT ? -1 : cur) : (int)(0u - (unsigned)cur); }
        else { price = -1; prev = MAXINT; have_prev = true; }
        int v = 0;
#pragma unroll
        for (int j = 0; j < RPL; ++j)
          if (VM_EX(j) && s.p[j] == price) v = wadd(v, s.q[j]);
        v = __reduce_add_sync(FULL, v);
        if (v < 0) v = 0;
        int clean = price != -1 ? v : 0;
        cum = wadd(cum, clean);
        if (lane == l) { lp[ch] = price; lv[ch] = v; cumv[ch] = price != -1 ? cum : 0; }
      }
    }
    if (lane < n) {
      if (P.raw) reinterpret_cast<int4*>(P.raw + (size_t)e * n * 4)[lane] = make_int4(lp[0], lp[1], lv[0], lv[1]);
      if (P.l2) reinterpret_cast<int4*>(P.l2 + (size_t)e * n * 4)[lane] = make_int4(lp[0], lv[0], lp[1], lv[1]);
      if (P.norm && have_mid) {
        const float tick = (float)P.tick;
        float ga = 0.f, gb = 0.f;
        if (lp[0] != -1) ga = __fdiv_rn(__fsub_rn((float)lp[0], mid), tick);
        if (lp[1] != -1) gb = __fdiv_rn(__fsub_rn(mid, (float)lp[1]), tick);
        float va = vm_log1p_f32((float)(lp[0] != -1 ? lv[0] : 0));
        float vb = vm_log1p_f32((float)(lp[1] != -1 ? lv[1] : 0));
        float ca = vm_log1p_f32((float)cumv[0]);
        float cb = vm_log1p_f32((float)cumv[1]);
        float2* o = reinterpret_cast<float2*>(P.norm + (size_t)e * n * 6 + lane * 6);
        o[0] = make_float2(ga, gb);
        o[1] = make_float2(va, vb);
        o[2] = make_float2(ca, cb);
      }
    }
  }
  if (P.image) {
    const int H = P.H, W = P.W;
    // per-row volume histogram (both channels) in shared scratch: [2][H]
    for (int i = lane; i < 2 * H; i += 32) scratch[i] = 0;
    __syncwarp();
    const int ba = best_price_out<RPL, false>(asks), bb = best_price_out<RPL, true>(bids);
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      if (VM_EX(j)) {
        if (ba != -1 && asks.p[j] != -1) {
          long long d = (long long)asks.p[j] - ba;
          if (d >= 0) { long long r = d / P.tick; if (r < H) atomicAdd(&scratch[(int)r], asks.q[j]); }
        }
        if (bb != -1 && bids.p[j] != -1) {
          long long d = (long long)bb - bids.p[j];
          if (d >= 0) { long long r = d / P.tick; if (r < H) atomicAdd(&scratch[H + (int)r], bids.q[j]); }
        }
      }
    }
    __syncwarp();
    for (int i = lane; i < 2 * H; i += 32) scratch[i] = bar_length(scratch[i], W);
    __syncwarp();
    if ((P.img_dtype & 0xff) == VITMARL_IMG_BF16_PATCHES) {
      // the raster written directly as the ViT's patch matrix [tokens = (H/p)(W/p), p*p*2] (token = py*(W/p)+px, feature =
      // (ph, pw, c)): the same bytes as the NHWC image, so the encoder's patchify pass (a 2 x 67 MB copy) disappears.
      // Chunks are walked in OUTPUT order: 32 lanes x 16 bytes stay one contiguous 512-byte store.
      const int ps = P.img_dtype >> 8, cpr = ps / 4, cpp = ps * cpr, tpr = W / ps;      // chunks per patch row / per patch, patches per image row
      uint4* img = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.image) + (size_t)e * H * W * 2);
      const int total = H * W / 4;
      // patch sizes and grids are powers of two in practice: shifts instead of three integer divisions per chunk
      const bool pow2 = !(cpr & (cpr - 1)) && !(ps & (ps - 1)) && !(tpr & (tpr - 1));
      const int s_cpr = __ffs(cpr) - 1, s_cpp = __ffs(cpp) - 1, s_tpr = __ffs(tpr) - 1;
      for (int c = lane; c < total; c += 32) {
        int t, ph, pw0, py, px;
        if (pow2) {
          t = c >> s_cpp; const int within = c & (cpp - 1);
          ph = within >> s_cpr; pw0 = (within & (cpr - 1)) * 4;
          py = t >> s_tpr; px = t & (tpr - 1);
        } else {
          t = c / cpp; const int within = c - t * cpp;
          ph = within / cpr; pw0 = (within - ph * cpr) * 4;
          py = t / tpr; px = t - py * tpr;
        }
        const int r = py * ps + ph, x0 = px * ps + pw0;
        const int la = scratch[r], lb = scratch[H + r];
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = ((x0 + k) < la ? 0x3F80u : 0u) | ((x0 + k) < lb ? 0x3F800000u : 0u);
        img[c] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else if (P.img_dtype == VITMARL_IMG_BF16) {
      // 16 bytes = 4 pixels x 2 channels x bf16
      uint4* img = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.image) + (size_t)e * H * W * 2);
      const int chunks_per_row = W / 4, total = H * chunks_per_row;
      for (int c = lane; c < total; c += 32) {
        int r = c / chunks_per_row, x0 = (c - r * chunks_per_row) * 4;
        int la = scratch[r], lb = scratch[H + r];
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = ((x0 + k) < la ? 0x3F80u : 0u) | ((x0 + k) < lb ? 0x3F800000u : 0u);
        img[c] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      // 16 bytes = 8 pixels x 2 channels x u8
      uint4* img = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(P.image) + (size_t)e * H * W * 2);
      const int chunks_per_row = W / 8, total = H * chunks_per_row;
      for (int c = lane; c < total; c += 32) {
        int r = c / chunks_per_row, x0 = (c - r * chunks_per_row) * 8;
        int la = scratch[r], lb = scratch[H + r];
        unsigned w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int x = x0 + 2 * k;
          w[k] = (x < la ? 1u : 0u) | (x < lb ? 0x100u : 0u) | ((x + 1) < la ? 0x10000u : 0u) | ((x + 1) < lb ? 0x1000000u : 0u);
        }
        img[c] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- fused reward reductions
// The integer trade reductions of both reward functions over ONE environment's trade log in shared memory -- the same
// arithmetic as trade_stats_kernel (env_glue.cu; vision_env.py:2076-2078, 2156-2163, 2191, mm_env.py:1906-1936), int32
// wrap-around like XLA.  out[8] = [sum qty, sum |qty|, c_rl, buyQuant, sellQuant, TradedVolume, inventory_delta, other |qty|].
__device__ __forceinline__ void trade_stats_from_slab(const int32_t* sm_tr, int T, int agent_id, int tick, int32_t* out, int lane) {
  unsigned s_q = 0, s_abs = 0, s_rev = 0, s_buy = 0, s_sell = 0, s_other = 0;
  for (int r = lane; r < T; r += 32) {
    const int4 a = *reinterpret_cast<const int4*>(sm_tr + r * 8), b = *reinterpret_cast<const int4*>(sm_tr + r * 8 + 4);
    const bool executed = a.x >= 0;                                  // JOBA:827
    const int price = executed ? a.x : 0, qty = executed ? a.y : 0, ptid = executed ? b.z : 0, atid = executed ? b.w : 0;
    const bool mine = agent_id == ptid || agent_id == atid;         // evaluated on the zeroed row, as the reference does
    const unsigned aq = qty < 0 ? 0u - (unsigned)qty : (unsigned)qty;
    if (mine) {
      s_q += (unsigned)qty; s_abs += aq; s_rev += (unsigned)(price / tick) * aq;
      if ((qty >= 0 && agent_id == ptid) || (qty < 0 && agent_id == atid)) s_buy += aq;
      if ((qty < 0 && agent_id == ptid) || (qty >= 0 && agent_id == atid)) s_sell += aq;
    } else {
      s_other += aq;
    }
  }
  s_q = __reduce_add_sync(FULL, s_q); s_abs = __reduce_add_sync(FULL, s_abs); s_rev = __reduce_add_sync(FULL, s_rev);
  s_buy = __reduce_add_sync(FULL, s_buy); s_sell = __reduce_add_sync(FULL, s_sell); s_other = __reduce_add_sync(FULL, s_other);
  if (lane == 0) {
    reinterpret_cast<int4*>(out)[0] = make_int4((int)s_q, (int)s_abs, (int)s_rev, (int)s_buy);
    reinterpret_cast<int4*>(out)[1] = make_int4((int)s_sell, (int)(s_buy + s_sell), (int)(s_buy - s_sell), (int)s_other);
  }
}

// ---------------------------------------------------------------- the kernel
template <int RPL, int MINB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB) lob_kernel(const LobParams P) {
  extern __shared__ __align__(16) int32_t smem[];
  __shared__ __align__(8) uint64_t bars[kWarpsPerCta];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * kWarpsPerCta + warp;
  if (e >= P.E) return;   // whole warp exits; no block-wide barriers are used below

  const int N = P.N, T = P.T, M = P.M;
  const int side_words = N * 6, trade_words = P.do_step ? T * 8 : 0;
  int32_t* sm_asks = smem + (size_t)warp * (2 * side_words + trade_words);
  int32_t* sm_bids = sm_asks + side_words;
  int32_t* sm_tr = sm_bids + side_words;
  const uint32_t bar = smem_u32(&bars[warp]);

  const int32_t* g_asks = P.asks_in + (size_t)e * side_words;
  const int32_t* g_bids = P.bids_in + (size_t)e * side_words;

  // ---- load: TMA 1-D bulk copies (one elected lane) --------------------------------------
  if (P.bulk_ok) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    if (lane == 0) {
      uint32_t bytes = 2u * side_words * 4u + ((P.do_step && P.trades_in) ? trade_words * 4u : 0u);
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(smem_u32(sm_asks), g_asks, side_words * 4, bar);
      bulk_g2s(smem_u32(sm_bids), g_bids, side_words * 4, bar);
      if (P.do_step && P.trades_in) bulk_g2s(smem_u32(sm_tr), P.trades_in + (size_t)e * trade_words, trade_words * 4, bar);
    }
  } else {
    warp_copy_g2s(sm_asks, g_asks, side_words, lane);
    warp_copy_g2s(sm_bids, g_bids, side_words, lane);
    if (P.do_step && P.trades_in) warp_copy_g2s(sm_tr, P.trades_in + (size_t)e * trade_words, trade_words, lane);
  }
  // overlapped with the copies: fresh trades (marl_env.py:377) and the first message block
  if (P.do_step && !P.trades_in) {
    int4* t4 = reinterpret_cast<int4*>(sm_tr);
    for (int i = lane; i < trade_words / 4; i += 32) t4[i] = make_int4(-1, -1, -1, -1);
  }
  const int4* g_msgs = reinterpret_cast<const int4*>(P.msgs + (size_t)e * M * 8);
  int4 m0 = make_int4(0, 0, 0, 0), m1 = m0;   // this lane's message of the current block of 32
  if (P.do_step && lane < M) { m0 = ld_nc_v4(g_msgs + 2 * lane); m1 = ld_nc_v4(g_msgs + 2 * lane + 1); }

  if (P.bulk_ok) mbar_wait(bar, 0);
  __syncwarp();

  Side<RPL> asks, bids;
  float mid = 0.f;
  bool have_mid = false;
#pragma unroll 1
  for (int pass = 0;; ++pass) {   // pass 1 only re-reads the register cache after the literal path ran
  regs_from_smem<RPL>(asks, sm_asks, N, lane);
  regs_from_smem<RPL>(bids, sm_bids, N, lane);
  refresh_best<RPL, false>(asks, N, lane);
  refresh_best<RPL, true>(bids, N, lane);
  if (pass == 1) break;
  bool slow = P.do_step && !slab_is_tidy(sm_asks, 2 * N, lane);   // this warp runs the literal path (from some message on)

  if (P.do_step) {
    // ---- message loop: blocks of 32 messages, one per lane -------------------------------
    const int keep0 = M - P.n_keep;   // first message index whose best bid/ask is reported
    int carry_a = -1, carry_b = -1;   // ffill carries across blocks
    int last_a = 0, last_b = 0;
    int tr_next = 0;                  // fast path: first free trade row (fresh log: everything is free)
    if (P.trades_in) {
      // a caller-supplied log is usable by the fast path when its free rows (column 4 == -1, Q6) are a suffix
      int ff = slow_first(T, lane, [&](int r) { return sm_tr[r * 8 + 4] == -1; });
      if (ff < 0) ff = T;
      bool ok = true;
      for (int r = ff + lane; r < T; r += 32) ok = ok && sm_tr[r * 8 + 4] == -1;
      if (!__all_sync(FULL, ok)) slow = true;
      tr_next = ff;
    }
    // per-lane outputs of the current block of 32 messages
    int oa_p = -1, oa_v = 0, ob_p = -1, ob_v = 0;
    // block epilogue: forward fill, output rows, next block of messages
    auto finish_block = [&](int base, int cnt) {
        const int gi = base + lane;   // global message index of this lane
        if (P.do_ffill) {
          // marl_env.py:685-711 on both tracks
          if (gi == 0) {
            if (oa_p == -1) { oa_p = P.last_ask_price[(size_t)e * P.last_stride]; oa_v = 0; }
            if (ob_p == -1) { ob_p = P.last_bid_price[(size_t)e * P.last_stride]; ob_v = 0; }
          }
          if (oa_p == -1) oa_v = 0;
          if (ob_p == -1) ob_v = 0;
          unsigned lower = (lane == 31) ? FULL : ((2u << lane) - 1u);
          unsigned va = __ballot_sync(FULL, lane < cnt && oa_p != -1) & lower;
          unsigned vb = __ballot_sync(FULL, lane < cnt && ob_p != -1) & lower;
          int sa = __shfl_sync(FULL, oa_p, va ? 31 - __clz(va) : 0);
          int sb = __shfl_sync(FULL, ob_p, vb ? 31 - __clz(vb) : 0);
          oa_p = va ? sa : carry_a;
          ob_p = vb ? sb : carry_b;
          carry_a = __shfl_sync(FULL, oa_p, cnt - 1);
          carry_b = __shfl_sync(FULL, ob_p, cnt - 1);
          last_a = carry_a; last_b = carry_b;
        }
        if (lane < cnt && gi >= keep0) {
          if (P.best_asks) reinterpret_cast<int2*>(P.best_asks + (size_t)e * P.n_keep * 2)[gi - keep0] = make_int2(oa_p, oa_v);
          if (P.best_bids) reinterpret_cast<int2*>(P.best_bids + (size_t)e * P.n_keep * 2)[gi - keep0] = make_int2(ob_p, ob_v);
        }
        const int nb = gi + 32;   // next block (not prefetched: 8 registers matter more than one exposed load per 32 messages)
        if (nb < M) { m0 = ld_nc_v4(g_msgs + 2 * nb); m1 = ld_nc_v4(g_msgs + 2 * nb + 1); }
      oa_p = -1; oa_v = 0; ob_p = -1; ob_v = 0;
    };
    int base = 0, i = 0;
    // ---- fast path: tidy book, register cache ----
    for (; base < M && !slow; base += 32) {
      const int cnt = min(32, M - base);
      const int code = decode_message(m0, m1);        // every lane decodes its own message once
      for (i = 0; i < cnt; ++i) {
        const int c = __shfl_sync(FULL, code, i);
        if (!(c & 8)) { slow = true; break; }         // could break tidiness: literal path from here to the end of the step
        process_message<RPL>(asks, bids, m0, m1, c & 7, i, sm_tr, tr_next, N, T, P.init_id, lane);
        if (lane == i) {
          oa_p = best_price_out<RPL, false>(asks); oa_v = asks.vol;
          ob_p = best_price_out<RPL, true>(bids);  ob_v = bids.vol;
        }
      }
      if (slow) break;
      finish_block(base, cnt);
    }
    // ---- literal path on the shared-memory slabs (the register cache is dead from here on) ----
    if (slow) {
      if (base >= M) i = 0;
      for (; base < M; base += 32) {
        const int cnt = min(32, M - base);
        for (; i < cnt; ++i) {
          __syncwarp();
          const int4 ma = make_int4(__shfl_sync(FULL, m0.x, i), __shfl_sync(FULL, m0.y, i), __shfl_sync(FULL, m0.z, i), __shfl_sync(FULL, m0.w, i));
          const int4 mb = make_int4(__shfl_sync(FULL, m1.x, i), __shfl_sync(FULL, m1.y, i), __shfl_sync(FULL, m1.z, i), __shfl_sync(FULL, m1.w, i));
          slow_message(sm_asks, sm_bids, sm_tr, N, T, P.init_id, ma, mb, lane);
          const int2 ba = slow_best(sm_asks, N, false, lane), bb = slow_best(sm_bids, N, true, lane);
          if (lane == i) { oa_p = ba.x == MAXINT ? -1 : ba.x; oa_v = ba.y; ob_p = bb.x; ob_v = bb.y; }
        }
        finish_block(base, cnt);
        i = 0;
      }
    }
    if (P.do_ffill) {
      mid = __fdiv_rn((float)wadd(last_b, last_a), 2.0f);   // marl_env.py:467 (int32 sum -> f32 -> /2)
      have_mid = true;
      if (lane == 0 && P.mid_out) P.mid_out[e] = mid;
      // world clock, marl_env.py:406,468,482: final_time = time columns of the LAST message; delta in float32, one rounding per
      // operation, in the reference's order ((ft0 + ft1/1e9) - t0) - t1/1e9
      if (lane == 0 && P.time_out && M > 0) {
        const int2 ft = *reinterpret_cast<const int2*>(P.msgs + ((size_t)e * M + (M - 1)) * 8 + 6);
        const int2 t = *reinterpret_cast<const int2*>(P.time_in + (size_t)e * 2);
        const float a = __fadd_rn((float)ft.x, __fdiv_rn((float)ft.y, 1e9f));
        P.delta_time[e] = __fsub_rn(__fsub_rn(a, (float)t.x), __fdiv_rn((float)t.y, 1e9f));
        *reinterpret_cast<int2*>(P.time_out + (size_t)e * 2) = ft;
      }
    }
    // ---- reward-function reductions over the step's trade log while it is still on chip ----
    if (P.n_stat > 0) {
      __syncwarp();
      for (int a = 0; a < P.n_stat; ++a) trade_stats_from_slab(sm_tr, T, P.stat_ids[a], P.stat_tick, P.stats + ((size_t)e * P.n_stat + a) * 8, lane);
    }
    // ---- store: the shared-memory slabs are already up to date -> TMA bulk stores ---------
    if (P.bulk_ok) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(P.asks_out + (size_t)e * side_words, smem_u32(sm_asks), side_words * 4);
        bulk_s2g(P.bids_out + (size_t)e * side_words, smem_u32(sm_bids), side_words * 4);
        if (P.trades_out) bulk_s2g(P.trades_out + (size_t)e * trade_words, smem_u32(sm_tr), trade_words * 4);
        bulk_commit();
      }
    } else {
      __syncwarp();
      warp_copy_s2g(P.asks_out + (size_t)e * side_words, sm_asks, side_words, lane);
      warp_copy_s2g(P.bids_out + (size_t)e * side_words, sm_bids, side_words, lane);
      if (P.trades_out) warp_copy_s2g(P.trades_out + (size_t)e * trade_words, sm_tr, trade_words, lane);
    }
  } else if (P.mid_in) {
    mid = P.mid_in[e];
    have_mid = true;
  }
  if (!(slow && P.do_step && P.do_render)) break;
  }

  if (!P.do_step && (P.best_asks || P.best_bids)) {   // get_best_bid_and_ask_inclQuants
    if (lane == 0) {
      if (P.best_asks) reinterpret_cast<int2*>(P.best_asks)[e] = make_int2(best_price_out<RPL, false>(asks), asks.vol);
      if (P.best_bids) reinterpret_cast<int2*>(P.best_bids)[e] = make_int2(best_price_out<RPL, true>(bids), bids.vol);
    }
  }

  if (P.do_render) {
    // scratch for the raster histogram: the trades slab is still being read by the bulk store,
    // so the render-only kernel uses the (idle) book staging area and the fused kernel waits.
    int32_t* scratch = sm_asks;
    if (P.do_step && P.image) {
      if (P.bulk_ok && lane == 0) bulk_wait_read0();
      __syncwarp();
    }
    render_env<RPL>(asks, bids, P, e, have_mid, mid, scratch, lane);
  }
  if (P.do_step && P.bulk_ok && lane == 0) bulk_wait_read0();   // smem must outlive the bulk reads
}

// ---------------------------------------------------------------- host-side launch
static int launch_lob(cudaStream_t stream, LobParams& P) {
  if (P.E == 0) return VITMARL_OK;
  if (P.E < 0 || P.N < 1 || P.N > 256 || !P.asks_in || !P.bids_in) return VITMARL_EINVAL;
  if (P.do_step) {
    if (P.T < 1 || P.T > 1024 || P.M < 0 || P.n_keep < 0 || !P.asks_out || !P.bids_out) return VITMARL_EINVAL;
    if (!P.trades_out && P.n_stat <= 0) return VITMARL_EINVAL;                // the trade log may stay on chip only when its reductions are asked for
    if (P.n_stat < 0 || P.n_stat > 4 || (P.n_stat > 0 && (!P.stats || (reinterpret_cast<uintptr_t>(P.stats) & 15) || P.stat_tick < 1))) return VITMARL_EINVAL;
    if (P.last_stride < 1) P.last_stride = 1;
    if (P.M > 0 && !P.msgs) return VITMARL_EINVAL;
    if (P.n_keep > P.M) P.n_keep = P.M;
    if (P.do_ffill && (P.M < 1 || !P.last_ask_price || !P.last_bid_price || P.n_keep != P.M)) return VITMARL_EINVAL;
    if (P.M > 0 && (reinterpret_cast<uintptr_t>(P.msgs) & 15)) return VITMARL_EINVAL;
    if (P.n_keep > 0 && ((P.best_asks && (reinterpret_cast<uintptr_t>(P.best_asks) & 7)) ||
                         (P.best_bids && (reinterpret_cast<uintptr_t>(P.best_bids) & 7)))) return VITMARL_EINVAL;
  }
  if (P.do_render) {
    if (P.n_levels < 1 || P.n_levels > 32 || P.tick < 1) return VITMARL_EINVAL;
    if ((P.raw && (reinterpret_cast<uintptr_t>(P.raw) & 15)) || (P.l2 && (reinterpret_cast<uintptr_t>(P.l2) & 15)) ||
        (P.norm && (reinterpret_cast<uintptr_t>(P.norm) & 7))) return VITMARL_EINVAL;
    if (P.image) {
      const int kind = P.img_dtype & 0xff, ps = P.img_dtype >> 8;
      if (kind == VITMARL_IMG_BF16_PATCHES) {
        if (ps < 4 || ps % 4 || P.H % ps || P.W % ps) return VITMARL_EINVAL;
      } else if (P.img_dtype != VITMARL_IMG_U8 && P.img_dtype != VITMARL_IMG_BF16) {
        return VITMARL_EINVAL;
      }
      if (P.H < 1 || P.W < 8 || (P.W % 8) || 2 * P.H > 12 * P.N) return VITMARL_EINVAL;
      if (reinterpret_cast<uintptr_t>(P.image) & 15) return VITMARL_EINVAL;
    }
  }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  P.bulk_ok = ((P.N * 24) % 16 == 0) && al16(P.asks_in) && al16(P.bids_in) &&
              (!P.do_step || (al16(P.asks_out) && al16(P.bids_out) && (!P.trades_out || al16(P.trades_out)) && (!P.trades_in || al16(P.trades_in))));
  const int rpl = (P.N + 31) / 32;
  const size_t smem = (size_t)kWarpsPerCta * (2 * P.N * 6 + (P.do_step ? P.T * 8 : 0)) * sizeof(int32_t);
  const dim3 grid((P.E + kWarpsPerCta - 1) / kWarpsPerCta), block(kWarpsPerCta * 32);
  cudaError_t err = cudaSuccess;
#define VM_LAUNCH(R, B)                                                                                   \
  do {                                                                                                    \
    if (smem > 48 * 1024) err = cudaFuncSetAttribute(lob_kernel<R, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (err == cudaSuccess) err = cudaFuncSetAttribute(lob_kernel<R, B>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  \
    if (err == cudaSuccess) lob_kernel<R, B><<<grid, block, smem, stream>>>(P);                           \
  } while (0)
  // MINB (resident CTAs per SM the register budget is sized for): 5 x 4 warps at 96 registers measured fastest for
  // N <= 128; 6 or 7 CTAs (80 / 72 registers) are slower although nothing spills inside the message loop.
  switch (rpl) {
    case 1: VM_LAUNCH(1, 5); break;
    case 2: VM_LAUNCH(2, 5); break;
    case 3: VM_LAUNCH(3, 5); break;
    case 4: VM_LAUNCH(4, 5); break;
    case 5: VM_LAUNCH(5, 3); break;
    case 6: VM_LAUNCH(6, 3); break;
    case 7: VM_LAUNCH(7, 2); break;
    default: VM_LAUNCH(8, 2); break;
  }
#undef VM_LAUNCH
  if (err == cudaSuccess) err = cudaGetLastError();
  return check_cuda(err);
}

}  // namespace vitmarl

using vitmarl::LobParams;

extern "C" int vitmarl_lob_step(void* stream, int E, int N, int T, int M, int n_keep,
                                const int32_t* asks_in, const int32_t* bids_in, const int32_t* trades_in,
                                const int32_t* msgs, int32_t* asks_out, int32_t* bids_out, int32_t* trades_out,
                                int32_t* best_asks, int32_t* best_bids, int cancel_mode, int32_t init_id) {
  if (cancel_mode != 0 && cancel_mode != 1) return VITMARL_EUNSUPPORTED;   // JOBA:130-163 need jax.random
  LobParams P{};
  P.E = E; P.N = N; P.T = T; P.M = M; P.n_keep = n_keep; P.init_id = init_id;
  P.asks_in = asks_in; P.bids_in = bids_in; P.trades_in = trades_in; P.msgs = msgs;
  P.asks_out = asks_out; P.bids_out = bids_out; P.trades_out = trades_out;
  P.best_asks = best_asks; P.best_bids = best_bids;
  P.do_step = 1;
  return vitmarl::launch_lob(static_cast<cudaStream_t>(stream), P);
}

extern "C" int vitmarl_lob_best_bid_ask(void* stream, int E, int N, const int32_t* asks, const int32_t* bids,
                                        int32_t* best_ask, int32_t* best_bid) {
  if (!best_ask || !best_bid) return VITMARL_EINVAL;
  LobParams P{};
  P.E = E; P.N = N; P.asks_in = asks; P.bids_in = bids; P.best_asks = best_ask; P.best_bids = best_bid;
  return vitmarl::launch_lob(static_cast<cudaStream_t>(stream), P);
}

extern "C" int vitmarl_lob_render(void* stream, int E, int N, int n_levels, int tick_size,
                                  const int32_t* asks, const int32_t* bids, const float* mid_price,
                                  int32_t* raw, int32_t* l2, float* norm, void* image, int img_dtype, int H, int W) {
  if (norm && !mid_price) return VITMARL_EINVAL;
  LobParams P{};
  P.E = E; P.N = N; P.asks_in = asks; P.bids_in = bids;
  P.do_render = 1; P.n_levels = n_levels; P.tick = tick_size; P.mid_in = mid_price;
  P.raw = raw; P.l2 = l2; P.norm = norm;
  P.image = (img_dtype == VITMARL_IMG_NONE) ? nullptr : image; P.img_dtype = img_dtype; P.H = H; P.W = W;
  return vitmarl::launch_lob(static_cast<cudaStream_t>(stream), P);
}

extern "C" int vitmarl_env_step2(void* stream, const VitmarlEnvStepArgs* a) {
  if (!a) return VITMARL_EINVAL;
  if (a->cancel_mode != 0 && a->cancel_mode != 1) return VITMARL_EUNSUPPORTED;
  LobParams P{};
  P.E = a->E; P.N = a->N; P.T = a->T; P.M = a->M; P.n_keep = a->M; P.init_id = a->init_id;
  P.asks_in = a->asks_in; P.bids_in = a->bids_in; P.msgs = a->msgs;
  P.asks_out = a->asks_out; P.bids_out = a->bids_out; P.trades_out = a->trades_out;
  P.best_asks = a->best_asks; P.best_bids = a->best_bids;
  P.do_step = 1; P.do_ffill = 1; P.last_ask_price = a->last_ask_price; P.last_bid_price = a->last_bid_price;
  P.last_stride = a->last_price_stride; P.mid_out = a->mid_price;
  const bool img = a->image && a->img_dtype != VITMARL_IMG_NONE;
  P.do_render = (a->raw || a->l2 || a->norm || img) ? 1 : 0;
  P.n_levels = a->n_levels; P.tick = a->tick_size; P.raw = a->raw; P.l2 = a->l2; P.norm = a->norm;
  P.image = img ? a->image : nullptr; P.img_dtype = a->img_dtype; P.H = a->H; P.W = a->W;
  P.n_stat = a->n_stat_agents; P.stats = a->trade_stats; P.stat_tick = a->tick_size;
  for (int i = 0; i < 4; ++i) P.stat_ids[i] = a->stat_agent_ids[i];
  if ((a->time_in != nullptr) != (a->time_out != nullptr) || (a->time_out != nullptr) != (a->delta_time != nullptr)) {
    vitmarl::set_last_error("env_step2: time_in / time_out / delta_time go together");
    return VITMARL_EINVAL;
  }
  P.time_in = a->time_in; P.time_out = a->time_out; P.delta_time = a->delta_time;
  return vitmarl::launch_lob(static_cast<cudaStream_t>(stream), P);
}

extern "C" int vitmarl_env_step(void* stream, int E, int N, int T, int M,
                                const int32_t* asks_in, const int32_t* bids_in, const int32_t* msgs,
                                const int32_t* last_ask_price, const int32_t* last_bid_price,
                                int32_t* asks_out, int32_t* bids_out, int32_t* trades_out,
                                int32_t* best_asks, int32_t* best_bids, float* mid_price,
                                int n_levels, int tick_size, int32_t* raw, int32_t* l2, float* norm,
                                void* image, int img_dtype, int H, int W, int cancel_mode, int32_t init_id) {
  if (cancel_mode != 0 && cancel_mode != 1) return VITMARL_EUNSUPPORTED;
  LobParams P{};
  P.E = E; P.N = N; P.T = T; P.M = M; P.n_keep = M; P.init_id = init_id;
  P.asks_in = asks_in; P.bids_in = bids_in; P.msgs = msgs;
  P.asks_out = asks_out; P.bids_out = bids_out; P.trades_out = trades_out;
  P.best_asks = best_asks; P.best_bids = best_bids;
  P.do_step = 1; P.do_ffill = 1; P.last_ask_price = last_ask_price; P.last_bid_price = last_bid_price; P.mid_out = mid_price;
  P.do_render = (raw || l2 || norm || (image && img_dtype != VITMARL_IMG_NONE)) ? 1 : 0;
  P.n_levels = n_levels; P.tick = tick_size; P.raw = raw; P.l2 = l2; P.norm = norm;
  P.image = (img_dtype == VITMARL_IMG_NONE) ? nullptr : image; P.img_dtype = img_dtype; P.H = H; P.W = W;
  return vitmarl::launch_lob(static_cast<cudaStream_t>(stream), P);
}
