"""Seeded synthetic LOBSTER-format inputs (host side, NumPy only).

The reference trains on LOBSTER CSVs that are not redistributable (and absent here), so
benchmarks and parity tests use synthetic books / message streams whose value ranges
follow the reference's own random generators (``gymnax_exchange/utils/utils.py:117-270``:
prices around 2.1-2.2 M, quantities 0-500, time deltas 0-2 s) and whose wire format is the
loader's ``[type, side, qty, price, order_id, trader_id, time_s, time_ns]`` int32 rows
(``gymnax_exchange/jaxlobster/lobster_loader.py:779-781``, ``jaxob_constants.py:76-83``).

Mixture per data message: 50 % passive limit, 15 % aggressive limit,
25 % partial cancel of a previously sent order, 5 % delete (type 3), 5 % cancel of an unknown
id (exercises the reference's cancel fall-backs).  Everything is vectorised over
environments; a stream is a pure function of ``(seed, E, config)``.
"""
from __future__ import annotations

import dataclasses

import numpy as np

__all__ = ["SynthConfig", "make_l2_books", "init_msgs_from_l2_batched", "MessageStream", "lobster_csv_rows"]


@dataclasses.dataclass(frozen=True)
class SynthConfig:
    mid: int = 2_200_000
    tick: int = 100
    n_levels: int = 10            # L2 depth of the initial book (book_depth, jaxob_config.py:166)
    max_gap_ticks: int = 3
    max_qty: int = 500
    p_passive: float = 0.50
    p_aggressive: float = 0.15
    p_partial_cancel: float = 0.25
    p_delete: float = 0.05        # remainder (0.05): cancel of an unknown id
    passive_depth_ticks: int = 10
    first_order_id: int = 1_000_000
    ring: int = 64                # remembered own orders per env (targets for cancels)
    start_time_s: int = 34_200    # jaxob_constants.py:21


def make_l2_books(E: int, seed: int = 1234, cfg: SynthConfig = SynthConfig()) -> np.ndarray:
    """-> int32 [E, 4*L] rows ``(ask_p, ask_q, bid_p, bid_q) x L`` like a LOBSTER orderbook row.
    Best ask = mid + tick, best bid = mid - tick, level gaps tick*U{1..3}, sizes U{1..500}."""
    rng = np.random.default_rng(seed)
    L = cfg.n_levels
    gaps_a = rng.integers(1, cfg.max_gap_ticks + 1, size=(E, L)) * cfg.tick
    gaps_b = rng.integers(1, cfg.max_gap_ticks + 1, size=(E, L)) * cfg.tick
    gaps_a[:, 0] = cfg.tick
    gaps_b[:, 0] = cfg.tick
    ask_p = cfg.mid + np.cumsum(gaps_a, axis=1)
    bid_p = cfg.mid - np.cumsum(gaps_b, axis=1)
    ask_q = rng.integers(1, cfg.max_qty + 1, size=(E, L))
    bid_q = rng.integers(1, cfg.max_qty + 1, size=(E, L))
    return np.stack((ask_p, ask_q, bid_p, bid_q), axis=2).reshape(E, 4 * L).astype(np.int32)


def init_msgs_from_l2_batched(book_l2: np.ndarray, time=(34_200, 0), init_id: int = -2) -> np.ndarray:
    """Batched ``init_msgs_from_l2`` (``JaxOrderBookArrays.py:913-942``, ``base_env.py:242-293``):
    2L limit messages per env; even rows ask(-1), odd rows bid(+1), order_id = init_id,
    trader_id = init_id - k, shared time.  -> int32 [E, 2L, 8]"""
    book_l2 = np.asarray(book_l2, dtype=np.int32)
    E = book_l2.shape[0]
    L = book_l2.shape[1] // 4
    data = book_l2.reshape(E, 2 * L, 2)
    out = np.zeros((E, 2 * L, 8), dtype=np.int32)
    out[:, :, 3] = data[:, :, 0]
    out[:, :, 2] = data[:, :, 1]
    out[:, :, 0] = 1
    out[:, 0::2, 1] = -1
    out[:, 1::2, 1] = 1
    out[:, :, 4] = init_id
    out[:, :, 5] = init_id - np.arange(2 * L, dtype=np.int32)[None, :]
    t = np.broadcast_to(np.asarray(time, dtype=np.int32), (E, 2))
    out[:, :, 6] = t[:, 0:1]
    out[:, :, 7] = t[:, 1:2]
    return out


class MessageStream:
    """Stateful generator of per-step message blocks ``[E, M, 8]`` (int32)."""

    def __init__(self, E: int, seed: int = 1234, cfg: SynthConfig = SynthConfig()):
        self.E, self.cfg = E, cfg
        self.rng = np.random.default_rng(seed + 7919)
        self.next_oid = np.full((E,), cfg.first_order_id, dtype=np.int64)
        self.t_ns = np.full((E,), cfg.start_time_s * 1_000_000_000, dtype=np.int64)
        R = cfg.ring
        self.ring_oid = np.zeros((E, R), dtype=np.int64)
        self.ring_side = np.zeros((E, R), dtype=np.int64)
        self.ring_price = np.zeros((E, R), dtype=np.int64)
        self.ring_qty = np.zeros((E, R), dtype=np.int64)
        self.ring_n = np.zeros((E,), dtype=np.int64)

    def _one(self) -> np.ndarray:
        c, rng, E = self.cfg, self.rng, self.E
        ar = np.arange(E)
        u = rng.random(E)
        side = rng.integers(0, 2, size=E) * 2 - 1                       # +1 bid / -1 ask
        self.t_ns += rng.integers(0, 2_000_000_001, size=E)
        ts, tns = self.t_ns // 1_000_000_000, self.t_ns % 1_000_000_000
        msg = np.zeros((E, 8), dtype=np.int64)
        msg[:, 6], msg[:, 7] = ts, tns

        k_pass = u < c.p_passive
        k_aggr = ~k_pass & (u < c.p_passive + c.p_aggressive)
        k_part = ~k_pass & ~k_aggr & (u < c.p_passive + c.p_aggressive + c.p_partial_cancel)
        k_del = ~k_pass & ~k_aggr & ~k_part & (u < c.p_passive + c.p_aggressive + c.p_partial_cancel + c.p_delete)
        k_unk = ~(k_pass | k_aggr | k_part | k_del)
        has = self.ring_n > 0
        # cancels with nothing to cancel yet become passive limits
        k_pass = k_pass | ((k_part | k_del) & ~has)
        k_part, k_del = k_part & has, k_del & has

        # limit orders (passive rest behind the touch; aggressive cross the opposite touch)
        depth = rng.integers(0, c.passive_depth_ticks + 1, size=E)
        through = rng.integers(1, c.max_gap_ticks + 1, size=E)
        p_pass = c.mid - side * c.tick * (1 + depth)
        p_aggr = c.mid + side * c.tick * through
        qty = rng.integers(1, c.max_qty + 1, size=E)
        qty_aggr = rng.integers(1, 2 * c.max_qty // 2 + 1, size=E)
        lim = k_pass | k_aggr
        oid = self.next_oid.copy()
        msg[lim, 0] = 1
        msg[lim, 1] = side[lim]
        msg[lim, 2] = np.where(k_aggr, qty_aggr, qty)[lim]
        msg[lim, 3] = np.where(k_aggr, p_aggr, p_pass)[lim]
        msg[lim, 4] = oid[lim]
        msg[lim, 5] = oid[lim]                                          # loader: trader_id = order_id
        self.next_oid += lim
        # remember passive orders as cancel targets
        slot = self.ring_n % c.ring
        w = k_pass
        self.ring_oid[ar[w], slot[w]] = oid[w]
        self.ring_side[ar[w], slot[w]] = side[w]
        self.ring_price[ar[w], slot[w]] = p_pass[w]
        self.ring_qty[ar[w], slot[w]] = qty[w]
        self.ring_n += w

        # cancels / deletes of a remembered order
        live = np.minimum(np.maximum(self.ring_n, 1), c.ring)
        pick = (rng.integers(0, 1 << 30, size=E) % live).astype(np.int64)
        can = k_part | k_del
        r_oid, r_side = self.ring_oid[ar, pick], self.ring_side[ar, pick]
        r_price, r_qty = self.ring_price[ar, pick], np.maximum(self.ring_qty[ar, pick], 1)
        part_q = 1 + (rng.integers(0, 1 << 30, size=E) % r_qty)
        msg[can, 0] = np.where(k_del, 3, 2)[can]
        msg[can, 1] = r_side[can]
        msg[can, 2] = np.where(k_del, r_qty, part_q)[can]
        msg[can, 3] = r_price[can]
        msg[can, 4] = r_oid[can]
        msg[can, 5] = r_oid[can]
        self.ring_qty[ar[k_part], pick[k_part]] -= part_q[k_part]

        # cancel of an unknown id: falls back to initial liquidity at that price or to Q5
        msg[k_unk, 0] = 2
        msg[k_unk, 1] = side[k_unk]
        msg[k_unk, 2] = rng.integers(1, 50, size=E)[k_unk]
        msg[k_unk, 3] = p_pass[k_unk]
        msg[k_unk, 4] = 9_000_000 + rng.integers(0, 1000, size=E)[k_unk]
        msg[k_unk, 5] = msg[k_unk, 4]
        return msg.astype(np.int32)

    def next(self, M: int) -> np.ndarray:
        return np.stack([self._one() for _ in range(M)], axis=1)


def lobster_csv_rows(msgs_env: np.ndarray):
    """One env's stream as LOBSTER message-file rows (time, type, order_id, size, price, direction),
    the 6-column on-disk format the reference loader ingests (lobster_loader.py:618-658)."""
    rows = []
    for t, s, q, p, oid, _tid, ts, tns in np.asarray(msgs_env).tolist():
        rows.append((f"{ts}.{tns:09d}", t, oid, q, p, s))
    return rows
