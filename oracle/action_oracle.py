"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

NumPy restatement of the two DEFAULT action -> message tables of the reference's agents:
  * ExecutionAgent._getActionMsgs_fixedQuant_complex  (gymnax_exchange/jaxen/vision_env.py:1046-1142; action_space
    "fixed_quants_complex" is the Execution_EnvironmentConfig default, jaxob_config.py:108)
  * MarketMakingAgent._getActionMsgs_spread_skew      (gymnax_exchange/jaxen/mm_env.py:1352-1491; action_space "spread_skew" is the
    MarketMaking_EnvironmentConfig default, jaxob_config.py:34; multiplier_type "tick" default, "spread" also restated)

Integer parts are exact restatements.  The floating-point parts follow JAX's dtype rules with x64 disabled: int32 operands are
converted to float32 and EVERY operation rounds to float32 (prices ~3e7 exceed float32's 2^24 integer range, so the rounding is part
of the reference's behaviour); `//` on floats is jax.numpy.floor_divide = `_float_divmod` (C fmod, subtract, divide, sign fix-up,
round-half-away); `.astype(int32)` / `jnp.array(..., dtype=int32)` truncate toward zero.
PARITY UNPINNED: JAX is not installed in this image, so these restatements could not be run against the reference; NumPy == CUDA
bit for bit is what the tests establish."""
import numpy as np

f32 = np.float32
i32 = np.int32

QUANT_ARRAY = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [2, 0, 0, 0], [0, 2, 0, 0],
                        [0, 0, 2, 0], [0, 0, 0, 2], [5, 0, 0, 0], [0, 5, 0, 0], [0, 0, 5, 0], [0, 0, 0, 5]], np.int32)   # vision_env.py:1098-1112


def _wrap(x: int) -> int:
    """int32 wrap-around of a Python integer (XLA integer arithmetic)."""
    return int(np.int64(x).astype(np.int32)) if -2**63 <= x < 2**63 else int(((x + 2**31) % 2**32) - 2**31)


def _ifloordiv(a: int, b: int) -> int:
    """jnp.floor_divide on int32 (ufuncs.py: lax.div truncates, minus one when the signs differ and the remainder is not zero)."""
    return a // b       # Python's // is the same floor division


def float_floor_divide(x, y):
    """jnp.floor_divide on float32 = _float_divmod(x, y)[0] (jax/_src/numpy/ufuncs.py): every step in float32."""
    x, y = f32(x), f32(y)
    mod = f32(np.fmod(x, y))
    div = f32(f32(x - mod) / y)
    ind = (mod != 0) and (np.sign(y) != np.sign(mod))
    if ind:
        div = f32(div - f32(1))
    r = np.floor(np.abs(div) + f32(0.5)).astype(np.float32)      # lax.round: half away from zero
    return f32(np.copysign(r, div))


def _to_i32(x) -> int:
    """convert_element_type float32 -> int32: truncation toward zero, saturating like XLA on the CPU/GPU backends."""
    x = float(x)
    if np.isnan(x):
        return 0
    return int(max(-2**31, min(2**31 - 1, int(x))))


def exec_action_msgs_fixed_quants_complex(action: int, best_ask_price: int, best_bid_price: int, is_sell_task: int,
                                          task_to_execute: int, quant_executed: int, time, trader_id: int, tick_size: int = 100,
                                          n_ticks_in_book: int = 1, fixed_quant_value: int = 10, time_delay_obs_act: int = 0,
                                          placeholder_order_id: int = -9) -> np.ndarray:
    """vision_env.py:1046-1142, one environment -> int32 [4, 8] (type, side, quant, price, order id, trader id, time_s, time_ns)."""
    tick = tick_size
    best_ask = _wrap(_ifloordiv(best_ask_price, tick) * tick)                                  # :1068
    best_bid = _wrap(_ifloordiv(best_bid_price, tick) * tick)                                  # :1069
    if not is_sell_task:                                                                       # buy_task_prices :1072-1078
        FT = best_ask
        M = _wrap(_ifloordiv(_ifloordiv(_wrap(best_bid + best_ask), 2), tick) * tick)
        NT = best_bid
        PP = _wrap(best_bid - tick * n_ticks_in_book)
    else:                                                                                      # sell_task_prices :1079-1086
        FT = best_bid
        half = f32(f32(_wrap(best_bid + best_ask)) / f32(2))                                   # int32 sum -> float32 true divide
        M = _to_i32(f32(np.ceil(float_floor_divide(half, tick)) * f32(tick)))
        NT = best_ask
        PP = _wrap(best_ask + tick * n_ticks_in_book)
    price_levels = [FT, M, NT, PP]
    quants = [_wrap(int(q) * fixed_quant_value) for q in QUANT_ARRAY[action]]                  # :1113
    side = 1 - int(bool(is_sell_task)) * 2                                                     # :1117
    quant_left = _wrap(task_to_execute - quant_executed)                                       # :1126
    total = _wrap(sum(quants))
    # :1128-1132  jnp.where(cond, int32 quants, jnp.floor(int32 array) -> float32).astype(int32): BOTH branches pass through float32
    if total <= quant_left:
        quants = [_to_i32(f32(q)) for q in quants]
    else:
        quants = [_to_i32(np.floor(f32(_wrap(int(q) * quant_left)))) for q in QUANT_ARRAY[1]]
    t = [_wrap(int(time[0]) + time_delay_obs_act), _wrap(int(time[1]) + time_delay_obs_act)]   # :1122-1125 (the delay is added to BOTH fields)
    out = np.zeros((4, 8), np.int32)
    for k in range(4):
        out[k] = (1, side, quants[k], price_levels[k], placeholder_order_id, trader_id, t[0], t[1])
    return out


def mm_action_msgs_spread_skew(action: int, best_ask_price: int, best_bid_price: int, time, trader_id: int, tick_size: int = 100,
                               spread_multiplier: float = 3.0, skew_multiplier: float = 5.0, multiplier_type: str = "tick",
                               fixed_quant_value: int = 10, time_delay_obs_act: int = 0, placeholder_order_id: int = -9) -> np.ndarray:
    """mm_env.py:1352-1491, one environment -> int32 [2, 8]: a bid and an ask limit order around a skewed mid price."""
    tick = tick_size
    best_ask = _wrap(_ifloordiv(best_ask_price, tick) * tick)                                  # :1368
    best_bid = _wrap(_ifloordiv(best_bid_price, tick) * tick)                                  # :1369
    mid_price = f32(f32(_wrap(best_ask + best_bid)) / f32(2))                                  # :1370
    current_spread = _wrap(best_ask - best_bid)                                                # :1379
    spread_type, skew_type = action // 3, action % 3                                           # :1383-1384
    mult = f32(1.0) if spread_type == 0 else f32(spread_multiplier)                            # :1389
    new_spread = f32(f32(current_spread) * mult)                                               # :1390
    skew_ticks = f32(-skew_multiplier) if skew_type == 0 else (f32(0) if skew_type == 1 else f32(skew_multiplier))   # :1394-1396
    if multiplier_type == "spread":                                                            # :1400-1403
        skewed_mid = f32(mid_price + f32(skew_ticks * new_spread))
    else:
        skewed_mid = f32(mid_price + f32(skew_ticks * f32(tick)))
    half_spread = float_floor_divide(new_spread, 2)                                            # :1410
    bid_price = f32(skewed_mid - half_spread)                                                  # :1411-1412
    ask_price = f32(skewed_mid + half_spread)
    bid_price = f32(float_floor_divide(bid_price, tick) * f32(tick))                           # :1438-1439
    ask_price = f32(float_floor_divide(ask_price, tick) * f32(tick))
    t = [_wrap(int(time[0]) + time_delay_obs_act), _wrap(int(time[1]) + time_delay_obs_act)]   # :1466
    out = np.zeros((2, 8), np.int32)
    out[0] = (1, 1, fixed_quant_value, _to_i32(bid_price), placeholder_order_id, trader_id, t[0], t[1])    # :1449-1461
    out[1] = (1, -1, fixed_quant_value, _to_i32(ask_price), placeholder_order_id, trader_id, t[0], t[1])
    return out
