/* CPU ORACLE (test infrastructure, NOT product code) -- plain C restatement of the
 * reference order-book path; batched over environments, OpenMP over envs so that it can
 * also serve as the timed CPU baseline ("port") in bench.py.
 *
 * Follows gymnax_exchange/jaxob/JaxOrderBookArrays.py (JOBA) of the reference line by
 * line; each function cites the lines it restates.  Pinned by golden vectors G1/G2 (see
 * oracle/lob_oracle.py header) through tests/test_oracle_golden.py, and cross-checked
 * against the NumPy restatement on random streams (tests/test_oracle_cross.py).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -fwrapv -ffp-contract=off)
 * int32 arithmetic wraps (-fwrapv) like XLA; no fast-math; no fma contraction so that
 * vm_log1p_f32 is bit-identical with the NumPy and CUDA versions.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXINT 2147483647
#define F 6   /* book row: price, qty, oid, tid, time_s, time_ns (jaxob_constants.py:36-42) */
#define TF 8  /* trade row (jaxob_constants.py:44-52) */

/* JOBA:85-90 (Q3) */
static void wipe(int32_t* s, int N) {
  for (int r = 0; r < N; ++r)
    if (s[r * F + 1] <= 0)
      for (int k = 0; k < F; ++k) s[r * F + k] = -1;
}

/* JOBA:62-83 (Q1, Q2) */
static void add_order(int32_t* s, int N, const int32_t* m, int32_t q) {
  int idx = -1;
  for (int r = 0; r < N && idx < 0; ++r)
    for (int k = 0; k < F; ++k)
      if (s[r * F + k] == -1) { idx = r; break; }
  if (idx < 0) idx = N - 1;
  int32_t* row = s + idx * F;
  row[0] = m[3]; row[1] = q > 0 ? q : 0; row[2] = m[4]; row[3] = m[5]; row[4] = m[6]; row[5] = m[7];
  wipe(s, N);
}

/* JOBA:93-138 (Q4, Q5); cancel modes 2/3 are rejected by the caller */
static void cancel_order(int32_t* s, int N, const int32_t* m, int32_t init_id) {
  int idx = -1;
  for (int r = 0; r < N; ++r) if (s[r * F + 2] == m[4]) { idx = r; break; }
  if (idx < 0)
    for (int r = 0; r < N; ++r)
      if (s[r * F + 0] == m[3] && s[r * F + 2] <= init_id && s[r * F + 1] >= m[2]) { idx = r; break; }
  if (idx < 0) idx = N - 1;
  s[idx * F + 1] = s[idx * F + 1] - m[2];
  wipe(s, N);
}

/* JOBA:240-267 (Q9); is_bid selects _get_top_bid_order_idx / _get_top_ask_order_idx */
static int top_idx(const int32_t* s, int N, int is_bid) {
  int32_t best;
  if (is_bid) {
    best = s[0];
    for (int r = 1; r < N; ++r) if (s[r * F] > best) best = s[r * F];
  } else {
    best = MAXINT;
    for (int r = 0; r < N; ++r) { int32_t p = s[r * F] == -1 ? MAXINT : s[r * F]; if (p < best) best = p; }
  }
  int32_t min_s = MAXINT;
  for (int r = 0; r < N; ++r) { int32_t t = s[r * F] == best ? s[r * F + 4] : MAXINT; if (t < min_s) min_s = t; }
  int32_t min_ns = MAXINT;
  for (int r = 0; r < N; ++r) {
    int32_t t = s[r * F] == best ? s[r * F + 4] : MAXINT;
    int32_t tn = t == min_s ? s[r * F + 5] : MAXINT;
    if (tn < min_ns) min_ns = tn;
  }
  for (int r = 0; r < N; ++r) {
    int32_t t = s[r * F] == best ? s[r * F + 4] : MAXINT;
    int32_t tn = t == min_s ? s[r * F + 5] : MAXINT;
    if (tn == min_ns) return r;
  }
  return N - 1; /* unreachable: first(...)=-1 wraps to last row */
}

/* JOBA:171-330 (Q6, Q7, Q11) */
static int32_t match(int32_t* s, int N, int is_bid_book, int32_t* tr, int T, const int32_t* m, int32_t qtm) {
  const int32_t price = m[3], side = m[1];
  int top = top_idx(s, N, is_bid_book);
  for (;;) {
    int32_t tp = s[top * F];
    int cross = is_bid_book ? (tp >= price) : (tp <= price);
    if (!(cross && qtm > 0 && tp != -1)) break;
    int32_t q_top = s[top * F + 1];
    int32_t d = q_top - qtm;
    int32_t newq = d > 0 ? d : 0;
    qtm = qtm - q_top;
    int e = -1;
    for (int r = 0; r < T; ++r) if (tr[r * TF + 4] == -1) { e = r; break; }
    if (e < 0) e = T - 1;
    int32_t* t = tr + e * TF;
    t[0] = tp; t[1] = -side * (q_top - newq); t[2] = s[top * F + 2]; t[3] = m[4];
    t[4] = m[6]; t[5] = m[7]; t[6] = s[top * F + 3]; t[7] = m[5];
    s[top * F + 1] = newq;
    wipe(s, N);
    top = top_idx(s, N, is_bid_book);
  }
  return qtm;
}

/* JOBA:833-898 (Q10) */
static void best_pair(const int32_t* s, int N, int is_bid, int32_t* out2) {
  int32_t best;
  if (is_bid) {
    best = s[0];
    for (int r = 1; r < N; ++r) if (s[r * F] > best) best = s[r * F];
  } else {
    best = MAXINT;
    for (int r = 0; r < N; ++r) { int32_t p = s[r * F] == -1 ? MAXINT : s[r * F]; if (p < best) best = p; }
    if (best == MAXINT) best = -1;
  }
  int32_t v = 0;
  for (int r = 0; r < N; ++r) if (s[r * F] == best) v += s[r * F + 1];
  out2[0] = best; out2[1] = v;
}

/* JOBA:617-661 */
static void process(int32_t* asks, int32_t* bids, int32_t* tr, int N, int T, const int32_t* m, int32_t init_id) {
  const int32_t t = m[0], s = m[1];
  int idx = (((s == -1 && t == 1) || (s == 1 && t == 4)) ? 0 : 0)
          + (((s == 1 && t == 1) || (s == -1 && t == 4)) ? 1 : 0)
          + ((s == -1 && (t == 2 || t == 3)) ? 2 : 0)
          + ((s == 1 && (t == 2 || t == 3)) ? 3 : 0)
          + ((s == 0 && t == 0) ? 4 : 0);
  switch (idx) {
    case 0: { int32_t q = match(bids, N, 1, tr, T, m, m[2]); add_order(asks, N, m, q); break; }
    case 1: { int32_t q = match(asks, N, 0, tr, T, m, m[2]); add_order(bids, N, m, q); break; }
    case 2: cancel_order(asks, N, m, init_id); break;
    case 3: cancel_order(bids, N, m, init_id); break;
    default: break;
  }
}

/* JOBA:720-752 batched ("vmapped") over E environments.
 * trades_in may be NULL (trades start at -1, marl_env.py:377).  best_* are [E, n_keep, 2]
 * (the last n_keep messages); either may be NULL (scan_through_entire_array, JOBA:665). */
int lob_oracle_step(int E, int N, int T, int M, int n_keep, int32_t init_id,
                    const int32_t* asks_in, const int32_t* bids_in, const int32_t* trades_in,
                    const int32_t* msgs, int32_t* asks_out, int32_t* bids_out, int32_t* trades_out,
                    int32_t* best_asks, int32_t* best_bids, int nthreads) {
  if (n_keep > M) n_keep = M;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    int32_t* a = asks_out + (size_t)e * N * F;
    int32_t* b = bids_out + (size_t)e * N * F;
    int32_t* t = trades_out + (size_t)e * T * TF;
    if (a != asks_in + (size_t)e * N * F) memcpy(a, asks_in + (size_t)e * N * F, sizeof(int32_t) * N * F);
    if (b != bids_in + (size_t)e * N * F) memcpy(b, bids_in + (size_t)e * N * F, sizeof(int32_t) * N * F);
    if (trades_in) { if (t != trades_in + (size_t)e * T * TF) memcpy(t, trades_in + (size_t)e * T * TF, sizeof(int32_t) * T * TF); }
    else for (int i = 0; i < T * TF; ++i) t[i] = -1;
    for (int i = 0; i < M; ++i) {
      process(a, b, t, N, T, msgs + ((size_t)e * M + i) * 8, init_id);
      int k = i - (M - n_keep);
      if (k >= 0) {
        if (best_asks) best_pair(a, N, 0, best_asks + ((size_t)e * n_keep + k) * 2);
        if (best_bids) best_pair(b, N, 1, best_bids + ((size_t)e * n_keep + k) * 2);
      }
    }
  }
  return 0;
}

/* JOBA:881-898 batched */
int lob_oracle_best(int E, int N, const int32_t* asks, const int32_t* bids, int32_t* best_ask, int32_t* best_bid) {
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    best_pair(asks + (size_t)e * N * F, N, 0, best_ask + (size_t)e * 2);
    best_pair(bids + (size_t)e * N * F, N, 1, best_bid + (size_t)e * 2);
  }
  return 0;
}

/* ---- stage 2 ------------------------------------------------------------------------ */
/* k-th distinct smallest key with fill; keys are produced by key(r). JOBA:1117-1134 (Q12) */
static void vision_raw_one(const int32_t* asks, const int32_t* bids, int N, int n, int32_t* raw /* [n,2,2] */) {
  for (int ch = 0; ch < 2; ++ch) {
    const int32_t* s = ch == 0 ? asks : bids;
    int have_prev = 0; int32_t prev = 0;
    for (int l = 0; l < n; ++l) {
      int found = 0; int32_t cur = 0;
      for (int r = 0; r < N; ++r) {
        int32_t p = s[r * F];
        int32_t key = ch == 0 ? (p == -1 ? MAXINT : p) : (int32_t)(0u - (uint32_t)p);
        if (have_prev && key <= prev) continue;
        if (!found || key < cur) { cur = key; found = 1; }
      }
      int32_t price;
      if (found) { prev = cur; have_prev = 1; price = ch == 0 ? (cur == MAXINT ? -1 : cur) : (int32_t)(0u - (uint32_t)cur); }
      else { price = -1; /* fill -1 (asks) / -(1) (bids); later levels stay fill */ have_prev = 1; prev = MAXINT; }
      int32_t v = 0;
      for (int r = 0; r < N; ++r) if (s[r * F] == price) v += s[r * F + 1];
      if (v < 0) v = 0;
      raw[(l * 2 + 0) * 2 + ch] = price;
      raw[(l * 2 + 1) * 2 + ch] = v;
    }
  }
}

/* deterministic log1p, see oracle/lob_oracle.py::log1p_f32 (identical operation sequence) */
static float vm_log1p_f32(float x) {
  static const double LN2_HI = 0x1.62e42fee00000p-1, LN2_LO = 0x1.a39ef35793c76p-33, SQRT2 = 0x1.6a09e667f3bcdp+0;
  static const double C[13] = {1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17,
                               1.0 / 19, 1.0 / 21, 1.0 / 23, 1.0 / 25, 1.0 / 27};
  double xd = (double)x;
  if (xd != xd || xd < -1.0) return NAN;
  if (xd == -1.0) return -INFINITY;
  if (isinf(xd)) return INFINITY;
  double y = 1.0 + xd;
  uint64_t bits; memcpy(&bits, &y, 8);
  int k = (int)((bits >> 52) & 0x7FF) - 1023;
  bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
  double m; memcpy(&m, &bits, 8);
  if (m > SQRT2) { m = m * 0.5; k += 1; }
  double f = m - 1.0;
  double s = f / (2.0 + f);
  double z = s * s;
  double p = C[12];
  for (int i = 11; i >= 0; --i) { double t = p * z; p = t + C[i]; }
  p = p * z;
  double s2 = 2.0 * s;
  double t2 = s2 * p;
  double logm = s2 + t2;
  double kd = (double)k;
  double a = kd * LN2_LO;
  double b = a + logm;
  double c = kd * LN2_HI;
  double r = c + b;
  return (float)r;
}

float lob_oracle_log1p_f32(float x) { return vm_log1p_f32(x); }

/* vision_env.py:2804-2854 */
static void vision_norm_one(const int32_t* raw, int n, float mid, float tick, float* out /* [n,3,2] */) {
  for (int ch = 0; ch < 2; ++ch) {
    int32_t cum = 0;
    for (int l = 0; l < n; ++l) {
      int32_t price = raw[(l * 2 + 0) * 2 + ch], vol = raw[(l * 2 + 1) * 2 + ch];
      int valid = price != -1;
      int32_t clean = valid ? vol : 0;
      cum += clean;
      float gap = 0.0f;
      if (valid) { float pf = (float)price; float d = ch == 0 ? pf - mid : mid - pf; gap = d / tick; }
      out[(l * 3 + 0) * 2 + ch] = gap;
      out[(l * 3 + 1) * 2 + ch] = vm_log1p_f32((float)clean);
      out[(l * 3 + 2) * 2 + ch] = vm_log1p_f32((float)(valid ? cum : 0));
    }
  }
}

static int bar_length(int32_t vol, int W) {
  uint32_t u = (uint32_t)(vol > 0 ? vol : 0) + 1u;
  int e = 31 - __builtin_clz(u);
  uint32_t frac2 = ((u << (31 - e)) >> 29) & 3u;
  int l4 = 4 * e + (int)frac2;
  int len = (l4 * W) >> 6;
  return len < W ? len : W;
}

/* docs/RENDER_SPEC.md (builder-defined raster).  img is uint8 [H,W,2]. */
static void image_one(const int32_t* asks, const int32_t* bids, int N, int H, int W, int tick, uint8_t* img, int32_t* volbuf) {
  memset(img, 0, (size_t)H * W * 2);
  for (int ch = 0; ch < 2; ++ch) {
    const int32_t* s = ch == 0 ? asks : bids;
    int32_t bp[2]; best_pair(s, N, ch, bp);
    if (bp[0] == -1) continue;
    for (int r = 0; r < H; ++r) volbuf[r] = 0;
    for (int r = 0; r < N; ++r) {
      int32_t p = s[r * F];
      if (p == -1) continue;
      int64_t d = ch == 0 ? (int64_t)p - bp[0] : (int64_t)bp[0] - p;
      if (d < 0) continue;
      int64_t row = d / tick;
      if (row < H) volbuf[row] += s[r * F + 1];
    }
    for (int r = 0; r < H; ++r) {
      int len = bar_length(volbuf[r], W);
      for (int x = 0; x < len; ++x) img[((size_t)r * W + x) * 2 + ch] = 1;
    }
  }
}

/* get_vision_L2_state + normalize_vision_obs (+ optional raster), batched.
 * raw [E,n,2,2] i32, norm [E,n,3,2] f32 (nullable), img [E,H,W,2] u8 (nullable). */
int lob_oracle_render(int E, int N, int n_levels, int tick, const int32_t* asks, const int32_t* bids,
                      const float* mid_price, int32_t* raw, float* norm, uint8_t* img, int H, int W, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int e = 0; e < E; ++e) {
    int32_t rawbuf[4 * 64];
    int32_t volbuf[1024];
    const int32_t* a = asks + (size_t)e * N * F;
    const int32_t* b = bids + (size_t)e * N * F;
    int32_t* r = raw ? raw + (size_t)e * n_levels * 4 : rawbuf;
    vision_raw_one(a, b, N, n_levels, r);
    if (norm) vision_norm_one(r, n_levels, mid_price[e], (float)tick, norm + (size_t)e * n_levels * 6);
    if (img) image_one(a, b, N, H, W, tick, img + (size_t)e * H * W * 2, volbuf);
  }
  return 0;
}

/* marl_env.py:392-393,467,685-711 : forward-fill + mid price, batched.
 * best_* [E,M,2] are updated in place; last_* [E] are the previous step's final prices. */
int lob_oracle_ffill_mid(int E, int M, int32_t* best_asks, int32_t* best_bids,
                         const int32_t* last_ask_price, const int32_t* last_bid_price, float* mid_price) {
  for (int e = 0; e < E; ++e) {
    for (int sd = 0; sd < 2; ++sd) {
      int32_t* pq = (sd == 0 ? best_asks : best_bids) + (size_t)e * M * 2;
      int32_t last = (sd == 0 ? last_ask_price : last_bid_price)[e];
      if (pq[0] == -1) { pq[0] = last; pq[1] = 0; }
      for (int i = 0; i < M; ++i) if (pq[i * 2] == -1) pq[i * 2 + 1] = 0;
      int32_t prev = -1;
      for (int i = 0; i < M; ++i) { if (pq[i * 2] != -1) prev = pq[i * 2]; pq[i * 2] = prev; }
    }
    int32_t s = best_bids[((size_t)e * M + M - 1) * 2] + best_asks[((size_t)e * M + M - 1) * 2];
    mid_price[e] = (float)s / 2.0f;
  }
  return 0;
}

int lob_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
