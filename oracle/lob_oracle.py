"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy restatement of the
reference order-book path, written line by line from the cited reference code.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this module.  The product path
(``vitmarl_b200``) never does: it fails loudly when the CUDA library is missing.

Parity pinning: this restatement is pinned by golden vector G1 (real LOBSTER AMZN
2012-06-21 rows stored in the reference notebook ``creating images.ipynb`` cells 3-4,
extracted by ``tests/golden/make_g1.py``) and by the worked example G2
(``gymnax_exchange/jaxob/jorderbook.py:296-302``).  The reference itself (JAX) cannot
be imported in this image, so matching-with-fills, cancels and the quirk ledger
 are "written line-by-line from the cited code" and
cross-checked NumPy <-> C <-> CUDA.

All ``JOBA:`` citations are ``gymnax_exchange/jaxob/JaxOrderBookArrays.py`` in the
reference.  Arrays are int32; arithmetic wraps like XLA int32.

JAX semantics relied upon:
  * ``jnp.where(mask, size=1, fill_value=-1)[0]`` -> first index (row-major) or -1
  * a dynamic index of -1 wraps to the last row
  * ``jnp.unique(x, size=n, fill_value=v)`` -> sorted unique, truncated / padded with v
  * int32 / python-int true division -> float32 (x64 disabled)
"""
from __future__ import annotations

import numpy as np

MAXINT = 2_147_483_647          # jaxob_constants.py:4-6 (enum oddly named _64_Bit_Signed)
INIT_ID = -2                    # jaxob_constants.py:9
EMPTY = -1                      # jaxob_constants.py:11


def _w32(x: int) -> int:
    """Wrap a python int to int32 two's complement (XLA int32 arithmetic)."""
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


def _first(mask: np.ndarray) -> int:
    """jnp.where(mask, size=1, fill_value=-1)[0][0]"""
    idx = np.flatnonzero(mask)
    return int(idx[0]) if idx.size else -1


# ----------------------------------------------------------------------------- core ops
def init_orderside(n_orders: int = 100) -> np.ndarray:
    """JOBA:901-911"""
    return np.full((n_orders, 6), -1, dtype=np.int32)


def _remove_zero_neg_quant(side: np.ndarray) -> np.ndarray:
    """JOBA:85-90 : every row with qty<=0 becomes all -1 (Q3)."""
    side[side[:, 1] <= 0, :] = -1
    return side


def add_order(side: np.ndarray, msg: dict) -> np.ndarray:
    """JOBA:62-83.  Q1: the empty row is the first row (row-major) in which ANY field is
    -1; Q2: none -> index -1 -> last row overwritten."""
    rows = np.nonzero(side == -1)[0]          # row-major order of 2-D where
    emptyidx = int(rows[0]) if rows.size else -1
    side[emptyidx, :] = [msg["price"], max(0, msg["quantity"]), msg["orderid"],
                         msg["traderid"], msg["time"], msg["time_ns"]]
    return _remove_zero_neg_quant(side)


def get_init_id_match(side: np.ndarray, msg: dict, init_id: int = INIT_ID) -> int:
    """JOBA:120-138 (cancel_mode 0/1: no randomness; modes 2/3 unsupported, Q4)."""
    m = (side[:, 0] == msg["price"]) & (side[:, 2] <= init_id) & (side[:, 1] >= msg["quantity"])
    return _first(m)


def cancel_order(side: np.ndarray, msg: dict, init_id: int = INIT_ID) -> np.ndarray:
    """JOBA:93-117.  Q5: no match -> idx -1 -> last row's qty decremented."""
    idx = _first(side[:, 2] == msg["orderid"])
    if idx == -1:
        idx = get_init_id_match(side, msg, init_id)
    side[idx, 1] = _w32(int(side[idx, 1]) - msg["quantity"])
    return _remove_zero_neg_quant(side)


def _get_top_bid_order_idx(side: np.ndarray) -> int:
    """JOBA:240-251"""
    max_price = side[:, 0].max()
    times = np.where(side[:, 0] == max_price, side[:, 4], MAXINT)
    min_s = times.min()
    times_ns = np.where(times == min_s, side[:, 5], MAXINT)
    min_ns = times_ns.min()
    return _first(times_ns == min_ns)


def _get_top_ask_order_idx(side: np.ndarray) -> int:
    """JOBA:254-267 (Q9)"""
    prices = np.where(side[:, 0] == -1, MAXINT, side[:, 0])
    min_price = prices.min()
    times = np.where(side[:, 0] == min_price, side[:, 4], MAXINT)
    min_s = times.min()
    times_ns = np.where(times == min_s, side[:, 5], MAXINT)
    min_ns = times_ns.min()
    return _first(times_ns == min_ns)


def _match_order(top: int, side: np.ndarray, qtm: int, trades: np.ndarray, msg: dict):
    """JOBA:171-219.  Q6: free trade slot = first row whose column 4 is -1; Q7: sign uses
    the raw message side; Q11: qtm may go negative."""
    q_top = int(side[top, 1])
    newquant = max(0, _w32(q_top - qtm))
    qtm = _w32(qtm - q_top)
    emptyidx = _first(trades[:, 4] == -1)
    trades[emptyidx, :] = [side[top, 0], _w32(-msg["side"] * _w32(q_top - newquant)), side[top, 2],
                           msg["orderid"], msg["time"], msg["time_ns"], side[top, 3], msg["traderid"]]
    side[top, 1] = newquant
    _remove_zero_neg_quant(side)
    return qtm


def _match_against_bid_orders(side, qtm, price, trades, msg):
    """JOBA:269-299"""
    top = _get_top_bid_order_idx(side)
    while (side[top, 0] >= price) and (qtm > 0) and (side[top, 0] != -1):
        qtm = _match_order(top, side, qtm, trades, msg)
        top = _get_top_bid_order_idx(side)
    return qtm


def _match_against_ask_orders(side, qtm, price, trades, msg):
    """JOBA:301-330"""
    top = _get_top_ask_order_idx(side)
    while (side[top, 0] <= price) and (qtm > 0) and (side[top, 0] != -1):
        qtm = _match_order(top, side, qtm, trades, msg)
        top = _get_top_ask_order_idx(side)
    return qtm


def _msg_dict(data) -> dict:
    """JOBA:638-645"""
    return {"side": int(data[1]), "type": int(data[0]), "price": int(data[3]), "quantity": int(data[2]),
            "orderid": int(data[4]), "traderid": int(data[5]), "time": int(data[6]), "time_ns": int(data[7])}


def branch_index(s: int, t: int) -> int:
    """JOBA:649-653 (Q8: anything outside the table -> 0 = ask_lim; only (0,0) is a no-op)."""
    return (int(((s == -1) and (t == 1)) or ((s == 1) and (t == 4))) * 0
            + int(((s == 1) and (t == 1)) or ((s == -1) and (t == 4))) * 1
            + int(((s == -1) and (t == 2)) or ((s == -1) and (t == 3))) * 2
            + int(((s == 1) and (t == 2)) or ((s == 1) and (t == 3))) * 3
            + int((s == 0) and (t == 0)) * 4)


def process_message(data, asks, bids, trades, init_id: int = INIT_ID):
    """JOBA:617-660 cond_type_side* (GENERAL_EXCHANGE mode); in-place on the arrays."""
    msg = _msg_dict(data)
    idx = branch_index(msg["side"], msg["type"])
    if idx == 0:      # ask_lim JOBA:417-453
        msg["quantity"] = _match_against_bid_orders(bids, msg["quantity"], msg["price"], trades, msg)
        add_order(asks, msg)
    elif idx == 1:    # bid_lim JOBA:356-391
        msg["quantity"] = _match_against_ask_orders(asks, msg["quantity"], msg["price"], trades, msg)
        add_order(bids, msg)
    elif idx == 2:    # ask_cancel JOBA:455-478
        cancel_order(asks, msg, init_id)
    elif idx == 3:    # bid_cancel JOBA:392-415
        cancel_order(bids, msg, init_id)
    # idx == 4: doNothing JOBA:334-355


# ----------------------------------------------------------------------------- read-outs
def get_volume_at_price(side: np.ndarray, price: int) -> int:
    """JOBA:833-844"""
    return _w32(int(np.where(side[:, 0] == price, side[:, 1], 0).astype(np.int64).sum()))


def get_best_ask(asks: np.ndarray) -> int:
    """JOBA:846-855"""
    m = int(np.where(asks[:, 0] == -1, MAXINT, asks[:, 0]).min())
    return -1 if m == MAXINT else m


def get_best_bid(bids: np.ndarray) -> int:
    """JOBA:857-865"""
    return int(bids[:, 0].max())


def get_best_bid_and_ask_inclQuants(asks, bids):
    """JOBA:881-898 (Q10: empty side -> [-1, -(#rows whose price is -1)])."""
    ba, bb = get_best_ask(asks), get_best_bid(bids)
    return (np.array([ba, get_volume_at_price(asks, ba)], dtype=np.int32),
            np.array([bb, get_volume_at_price(bids, bb)], dtype=np.int32))


def scan_through_entire_array(msgs, book_state, init_id: int = INIT_ID):
    """JOBA:665-685"""
    asks, bids, trades = (np.array(a, dtype=np.int32, copy=True) for a in book_state)
    for m in np.asarray(msgs):
        process_message(m, asks, bids, trades, init_id)
    return asks, bids, trades


def scan_through_entire_array_save_bidask(msgs, book_state, n_steps: int, init_id: int = INIT_ID):
    """JOBA:720-752 -> ((asks,bids,trades),(best_asks[-n_steps:],best_bids[-n_steps:]))"""
    asks, bids, trades = (np.array(a, dtype=np.int32, copy=True) for a in book_state)
    msgs = np.asarray(msgs)
    b_asks = np.zeros((msgs.shape[0], 2), dtype=np.int32)
    b_bids = np.zeros((msgs.shape[0], 2), dtype=np.int32)
    for i, m in enumerate(msgs):
        process_message(m, asks, bids, trades, init_id)
        b_asks[i], b_bids[i] = get_best_bid_and_ask_inclQuants(asks, bids)
    return (asks, bids, trades), (b_asks[-n_steps:], b_bids[-n_steps:])


def init_msgs_from_l2(book_l2, time=None, init_id: int = INIT_ID) -> np.ndarray:
    """JOBA:913-942 / base_env.py:242-293 (Q13: oid=-2 for all, tid=-2-k)."""
    book_l2 = np.asarray(book_l2, dtype=np.int64)
    levels = book_l2.shape[0] // 4
    data = book_l2.reshape(levels * 2, 2)
    if time is None:
        time = (34200, 0)
    out = np.zeros((levels * 2, 8), dtype=np.int32)
    out[:, 3] = data[:, 0]
    out[:, 2] = data[:, 1]
    out[:, 0] = 1
    out[0::2, 1] = -1
    out[1::2, 1] = 1
    out[:, 4] = init_id
    out[:, 5] = init_id - np.arange(levels * 2)
    out[:, 6] = time[0]
    out[:, 7] = time[1]
    return out


def _unique_size(x: np.ndarray, size: int, fill_value: int) -> np.ndarray:
    """jnp.unique(x, size=size, fill_value=fill_value)"""
    u = np.unique(x)
    out = np.full((size,), fill_value, dtype=np.int32)
    k = min(size, u.size)
    out[:k] = u[:k]
    return out


def _level_prices(asks, bids, n_levels):
    """JOBA:1117-1126 (shared by get_L2_state :1088-1097)."""
    neg = (-bids[:, 0].astype(np.int64)).astype(np.int32)            # int32 wrap like XLA
    bid_prices = (-_unique_size(neg, n_levels, 1).astype(np.int64)).astype(np.int32)
    ask_prices = _unique_size(np.where(asks[:, 0] == -1, MAXINT, asks[:, 0]).astype(np.int32), n_levels, -1)
    ask_prices = np.where(ask_prices == MAXINT, -1, ask_prices).astype(np.int32)
    return ask_prices, bid_prices


def get_vision_L2_state(asks, bids, n_levels: int) -> np.ndarray:
    """JOBA:1108-1140 -> int32 [n_levels, 2 (price, vol), 2 (ask, bid)]  (Q12)"""
    ask_p, bid_p = _level_prices(asks, bids, n_levels)
    ask_v = np.array([max(0, get_volume_at_price(asks, int(p))) for p in ask_p], dtype=np.int32)
    bid_v = np.array([max(0, get_volume_at_price(bids, int(p))) for p in bid_p], dtype=np.int32)
    ask_raw = np.stack((ask_p, ask_v), axis=1)
    bid_raw = np.stack((bid_p, bid_v), axis=1)
    return np.stack((ask_raw, bid_raw), axis=2).astype(np.int32)


def get_L2_state(asks, bids, n_levels: int) -> np.ndarray:
    """JOBA:1075-1106 -> int32 [4*n_levels] = (ask_p, ask_q, bid_p, bid_q) per level"""
    v = get_vision_L2_state(asks, bids, n_levels)         # same prices / clamped volumes
    return np.stack((v[:, 0, 0], v[:, 1, 0], v[:, 0, 1], v[:, 1, 1]), axis=1).reshape(-1).astype(np.int32)


# ----------------------------------------------------------------------------- env glue on the path
def ffill_best_prices(prices_quants: np.ndarray, last_valid_price: int) -> np.ndarray:
    """marl_env.py:685-711 `_ffill_best_prices`."""
    pq = np.array(prices_quants, dtype=np.int32, copy=True)
    if pq[0, 0] == -1:
        pq[0, 0:2] = [last_valid_price, 0]
    pq[:, 1] = np.where(pq[:, 0] == -1, 0, pq[:, 1])
    prev = -1
    for i in range(pq.shape[0]):
        if pq[i, 0] != -1:
            prev = pq[i, 0]
        pq[i, 0] = prev
    return pq


def mid_price_f32(best_bid_price: int, best_ask_price: int) -> np.float32:
    """marl_env.py:190,467 : int32 sum (wrapping) -> float32 -> /2."""
    s = np.float32(_w32(int(best_bid_price) + int(best_ask_price)))
    return np.float32(s / np.float32(2.0))


# log1p: evaluated in float64 with a FIXED sequence of IEEE mul/add/div (no fma, no libm)
# so that NumPy, C and CUDA agree bit-for-bit; the float32 rounding of the float64
# result is (up to double-rounding ties) the correctly rounded log1p.  XLA-CPU's own
# float32 log1p polynomial is within 2 ulp(fp32) of this (documented tolerance vs JAX).
_LN2_HI = float.fromhex("0x1.62e42fee00000p-1")
_LN2_LO = float.fromhex("0x1.a39ef35793c76p-33")
_SQRT2 = float.fromhex("0x1.6a09e667f3bcdp+0")
_LOG_COEF = [1.0 / (2 * k + 1) for k in range(1, 14)]   # 1/3 .. 1/27  (correctly rounded python floats)


def log1p_f32(x) -> np.float32:
    """float32 log1p of an (integer-valued) float32, see note above.
    vision_env.py:2826,2831,2843,2847 call jnp.log1p on int32 -> float32."""
    x = np.float32(x)
    xd = float(x)
    if xd != xd or xd < -1.0:
        return np.float32(np.nan)
    if xd == -1.0:
        return np.float32(-np.inf)
    if xd == np.inf:
        return np.float32(np.inf)
    y = 1.0 + xd                                  # exact for integer-valued |x| < 2^53
    bits = np.float64(y).view(np.int64)
    k = int((bits >> 52) & 0x7FF) - 1023
    m = np.int64((int(bits) & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000).view(np.float64)
    m = float(m)
    if m > _SQRT2:
        m = m * 0.5
        k += 1
    f = m - 1.0
    s = f / (2.0 + f)
    z = s * s
    p = _LOG_COEF[-1]
    for c in reversed(_LOG_COEF[:-1]):
        p = p * z + c                             # python floats: separate mul, add (no fma)
    p = p * z                                     # z/3 + z^2/5 + ...
    logm = 2.0 * s + (2.0 * s) * p
    kd = float(k)
    r = kd * _LN2_HI + (kd * _LN2_LO + logm)
    return np.float32(r)


def normalize_vision_obs(raw: np.ndarray, mid_price, tick_size: int = 100) -> np.ndarray:
    """vision_env.py:2804-2854 -> float32 [n, 3 (gap, log vol, log cum vol), 2 (ask, bid)]"""
    raw = np.asarray(raw, dtype=np.int32)
    mid = np.float32(mid_price)
    tick = np.float32(tick_size)
    n = raw.shape[0]
    out = np.zeros((n, 3, 2), dtype=np.float32)
    for ch, sign in ((0, 1.0), (1, -1.0)):
        prices = raw[:, 0, ch]
        vols = raw[:, 1, ch]
        valid = prices != -1
        clean = np.where(valid, vols, 0).astype(np.int32)
        cum = np.cumsum(clean.astype(np.int64)).astype(np.int32)       # int32 wrap
        cum = np.where(valid, cum, 0).astype(np.int32)
        for i in range(n):
            if valid[i]:
                pf = np.float32(prices[i])
                d = np.float32(pf - mid) if ch == 0 else np.float32(mid - pf)
                out[i, 0, ch] = np.float32(d / tick)
            out[i, 1, ch] = log1p_f32(np.float32(clean[i]))
            out[i, 2, ch] = log1p_f32(np.float32(cum[i]))
    return out


# ----------------------------------------------------------------------------- builder-defined raster
def bar_length(vol: int, width: int) -> int:
    """docs/RENDER_SPEC.md: quarter-octave integer log2 thermometer length."""
    u = (max(0, int(vol)) + 1) & 0xFFFFFFFF
    e = u.bit_length() - 1
    frac2 = ((u << (31 - e)) >> 29) & 3
    l4 = 4 * e + frac2
    return min(width, (l4 * width) >> 6)


def render_image(asks, bids, height: int, width: int, tick_size: int = 100) -> np.ndarray:
    """docs/RENDER_SPEC.md (builder-defined; the reference renders no image).
    -> uint8 [H, W, 2] in {0,1}; channel 0 ask (rows ascend from best ask), 1 bid."""
    img = np.zeros((height, width, 2), dtype=np.uint8)
    ba, bb = get_best_ask(asks), get_best_bid(bids)
    for ch, side, best in ((0, asks, ba), (1, bids, bb)):
        if best == -1:
            continue
        vol = np.zeros((height,), dtype=np.int64)
        for p, q in zip(side[:, 0].tolist(), side[:, 1].tolist()):
            if p == -1:
                continue
            d = (p - best) if ch == 0 else (best - p)
            if d < 0:
                continue
            r = d // tick_size
            if r < height:
                vol[r] += q
        for r in range(height):
            v = _w32(int(vol[r]))
            img[r, :bar_length(v, width), ch] = 1
    return img


# ----------------------------------------------------------------------------- callers either side of the book step
def getCancelMsgs(bookside, agentID, size, side, cancel_time, cancel_time_ns) -> np.ndarray:
    """JOBA:756-782: cancel messages for the first `size` orders of a trader; fill -> the appended zero row."""
    book = np.concatenate([np.asarray(bookside, dtype=np.int32), np.zeros((1, 6), dtype=np.int32)], axis=0)
    idx = np.flatnonzero(book[:, 3] == agentID)[:size]
    idx = np.concatenate([idx, np.full((size - idx.size,), -1, dtype=np.int64)])
    out = np.zeros((size, 8), dtype=np.int32)
    out[:, 0] = 2
    out[:, 1] = side
    out[:, 2] = book[idx, 1]
    out[:, 3] = book[idx, 0]
    out[:, 4] = book[idx, 2]
    out[:, 5] = book[idx, 3]
    out[:, 6] = cancel_time
    out[:, 7] = cancel_time_ns
    return out


def get_agent_trades(trades, agent_id) -> np.ndarray:
    """JOBA:824-831"""
    trades = np.asarray(trades, dtype=np.int32)
    executed = np.where((trades[:, 0] >= 0)[:, None], trades, 0)
    mask2 = (agent_id == executed[:, 6]) | (agent_id == executed[:, 7])
    return np.where(mask2[:, None], executed, 0).astype(np.int32)


def filter_messages(action_msgs, cnl_msgs):
    """`_filter_messages` (vision_env.py:622-684; the same text in mm_env.py:509-571 and exec_env.py:611-673), literal:
    cancellations are netted against new orders at the same price.  jnp.where(mask, size=n, fill_value=-1) = ascending
    indices padded with -1; rank_rev(mask) = stable descending rank; int32 arithmetic."""
    action_msgs = np.array(action_msgs, dtype=np.int32)
    cnl_msgs = np.array(cnl_msgs, dtype=np.int32)
    n = action_msgs.shape[0]
    assert cnl_msgs.shape[0] == n, "(c >= a) * a needs as many cancel rows as action rows"

    def argsort_rev(arr):
        return (arr.shape[0] - 1 - np.argsort(arr[::-1], kind="stable"))[::-1]

    def rank_rev(arr):
        return np.argsort(argsort_rev(arr), kind="stable")

    def where_size(mask):
        idx = np.flatnonzero(mask)
        return np.concatenate([idx, np.full(n - idx.size, -1, dtype=np.int64)])

    pa, pc = action_msgs[:, 3], cnl_msgs[:, 3]
    res = (pc[None, :] == pa[:, None]) & (pa[:, None] != 0)
    a_mask, c_mask = res.any(axis=1), res.any(axis=0)
    a_i = where_size(a_mask)
    a = np.where(a_i == -1, 0, action_msgs[a_i][:, 2]).astype(np.int32)
    c_i = where_size(c_mask)
    c = np.where(c_i == -1, 0, cnl_msgs[c_i][:, 2]).astype(np.int32)
    rel = ((c >= a) * a).astype(np.int32)
    with np.errstate(over="ignore"):
        action_msgs[:, 2] = action_msgs[:, 2] - rel[rank_rev(a_mask)]
        action_msgs = np.where((action_msgs[:, 2] == 0)[:, None], 0, action_msgs).astype(np.int32)
        cnl_msgs[:, 2] = cnl_msgs[:, 2] - rel[rank_rev(c_mask)]
    return action_msgs, cnl_msgs


def agent_trade_stats(trades, agent_id, tick_size) -> np.ndarray:
    """The trade reductions the reward functions take of one environment's trades [T,8], int32 wrap-around like XLA:
    vision_env.py:2076-2077 (signed sum), :2156-2163 (agentQuant), :2191 (c_rl); mm_env.py:1906-1933 (buyQuant,
    sellQuant, TradedVolume, inventory_delta); last entry: sum |qty| of otherTrades (mm_env.py:1914)."""
    trades = np.asarray(trades, dtype=np.int32)
    executed = np.where((trades[:, 0] >= 0)[:, None], trades, 0).astype(np.int32)
    mask2 = (agent_id == executed[:, 6]) | (agent_id == executed[:, 7])
    agent = np.where(mask2[:, None], executed, 0).astype(np.int32)
    other = np.where(mask2[:, None], 0, executed).astype(np.int32)
    with np.errstate(over="ignore"):
        absq = np.abs(agent[:, 1])                       # int32: abs(INT_MIN) stays INT_MIN, as jnp.abs
        mask_buy = ((agent[:, 1] >= 0) & (agent_id == agent[:, 6])) | ((agent[:, 1] < 0) & (agent_id == agent[:, 7]))
        mask_sell = ((agent[:, 1] < 0) & (agent_id == agent[:, 6])) | ((agent[:, 1] >= 0) & (agent_id == agent[:, 7]))
        buy = np.where(mask_buy, absq, 0).astype(np.int32).sum(dtype=np.int32)
        sell = np.where(mask_sell, absq, 0).astype(np.int32).sum(dtype=np.int32)
        out = np.array([agent[:, 1].sum(dtype=np.int32), absq.sum(dtype=np.int32),
                        ((agent[:, 0] // np.int32(tick_size)) * absq).astype(np.int32).sum(dtype=np.int32),
                        buy, sell, np.int32(buy + sell), np.int32(buy - sell),
                        np.abs(other[:, 1]).sum(dtype=np.int32)], dtype=np.int32)
    return out


def get_data_messages(message_data, start, step_counter, n_data_msg_per_step, end_time_s=None) -> np.ndarray:
    """base_env.py:341-371 (`_get_data_messages`); lax.dynamic_slice_in_dim clamps the start index.
    end_time_s is given for ep_type == 'fixed_time' only."""
    message_data = np.asarray(message_data, dtype=np.int32)
    off = int(start) + n_data_msg_per_step * int(step_counter)
    off = max(0, min(off, message_data.shape[0] - n_data_msg_per_step))
    msgs = message_data[off:off + n_data_msg_per_step].copy()
    if end_time_s is not None:
        late = msgs[:, -2] >= end_time_s
        msgs[late, :-2] = 0
    return msgs


def build_step_msgs(message_data, start, step_counter, n_data, cancel_msgs, action_msgs, order_id_counter, perm=None,
                    end_time_s=None):
    """marl_env.py:272-344: data messages, order-id renumbering of the action messages (:314-319), shuffle (:322-324,
    with the permutation indices supplied by the caller's jax.random.permutation), concatenation (:344)."""
    action = np.array(action_msgs, dtype=np.int32, copy=True)
    Ma = action.shape[0]
    action[:, 4] = np.array([_w32(int(order_id_counter) - k) for k in range(Ma)], dtype=np.int32)
    if perm is not None:
        action = action[np.asarray(perm)]
    data = get_data_messages(message_data, start, step_counter, n_data, end_time_s)
    combined = np.concatenate([np.asarray(cancel_msgs, dtype=np.int32).reshape(-1, 8), action, data], axis=0)
    return combined, _w32(int(order_id_counter) - Ma)


def auto_reset(done, window, init_asks, init_bids, init_best_asks, init_best_bids, asks, bids, trades, best_asks, best_bids, mid,
               init_trades=None):
    """marl_env.py:737-766 (jax.lax.select(done, reset_state, stepped_state) per leaf) with the reset state of
    base_env.py:215-231 (index_tree(init_states_array, idx_data_window)) and marl_env.py:186-190 (best bid / ask tiled over the
    message slots, mid_price = float32((best_bid[0] + best_ask[0]) / 2) with x64 disabled).  Returns new copies."""
    asks, bids, trades = np.array(asks, np.int32), np.array(bids, np.int32), np.array(trades, np.int32)
    best_asks, best_bids, mid = np.array(best_asks, np.int32), np.array(best_bids, np.int32), np.array(mid, np.float32)
    M = best_asks.shape[1]
    for e in range(asks.shape[0]):
        if not done[e]:
            continue
        w = int(window[e])
        asks[e], bids[e] = init_asks[w], init_bids[w]
        trades[e] = -1 if init_trades is None else init_trades[w]
        best_asks[e] = np.tile(np.asarray(init_best_asks[w], np.int32)[None, :], (M, 1))
        best_bids[e] = np.tile(np.asarray(init_best_bids[w], np.int32)[None, :], (M, 1))
        ssum = (int(init_best_bids[w][0]) + int(init_best_asks[w][0]) + 2**31) % 2**32 - 2**31     # int32 wrap-around (XLA semantics)
        mid[e] = np.float32(np.float32(ssum) / np.float32(2))
    return asks, bids, trades, best_asks, best_bids, mid


def world_time_update(msgs: np.ndarray, time: np.ndarray):
    """World clock of MARLEnv.step_env (marl_env.py:406, 468, 482), one environment:
        final_time = combined_msgs[-1, -2:]
        new_delta_time = final_time[0] + final_time[1]/1e9 - time[0] - time[1]/1e9
    with x64 disabled: the int32 operands are converted to float32, every division / addition / subtraction rounds to float32,
    evaluated left to right as Python parses it: ((ft0 + ft1/1e9) - t0) - t1/1e9.  -> (final_time int32 [2], delta float32)."""
    ft = np.array(msgs[-1, 6:8], np.int32)
    f32 = np.float32
    a = f32(f32(ft[0]) + f32(f32(ft[1]) / f32(1e9)))
    b = f32(a - f32(time[0]))
    d = f32(b - f32(f32(time[1]) / f32(1e9)))
    return ft, d
