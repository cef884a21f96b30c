"""Shared generators for the parity tests (test infrastructure)."""
import numpy as np


def lobster_to_msg(row):
    """LOBSTER (time, type, order_id, size, price, direction) -> loader message
    [type, side, qty, price, trader_id(=order_id), order_id, t_s, t_ns]
    (gymnax_exchange/jaxlobster/lobster_loader.py:618-658, 779-781)."""
    tm, ty, oid, sz, pr, d = row
    s, ns = str(tm).split(".")
    ns = int((ns + "0" * 9)[:9])
    ty, d = int(ty), int(d)
    if ty == 4:
        d, ty = -d, 1
    if ty == 3:
        ty = 2
    return np.array([ty, d, int(sz), int(pr), int(oid), int(oid), int(s), ns], dtype=np.int32)


def fuzz_case(rng, E, N, T, M, fill=0.6, weird=0.15, tidy=False):
    """Adversarial small-range books + messages: collisions of ids / prices / times, -1 fields,
    negative quantities, unknown (type, side) pairs, full books and full trade logs.
    tidy=True keeps the collisions but makes every resting row well formed (no -1 field, qty > 0), every limit
    message restable and the pre-filled trade rows a prefix: the state the CUDA kernel's register fast path
    handles (a cancel with a negative quantity still shows up now and then and sends an environment to the
    literal path in the middle of its step)."""
    if tidy:
        weird = 0.0
    asks = np.full((E, N, 6), -1, dtype=np.int32)
    bids = np.full((E, N, 6), -1, dtype=np.int32)
    for side, lo, hi in ((asks, 105, 112), (bids, 98, 106)):
        live = rng.random((E, N)) < fill
        side[..., 0] = rng.integers(lo, hi, size=(E, N))
        side[..., 1] = rng.integers(1, 30, size=(E, N))
        side[..., 2] = rng.integers(-6, 12, size=(E, N))
        side[..., 3] = rng.integers(-4, 4, size=(E, N))
        side[..., 4] = rng.integers(0, 4, size=(E, N))
        side[..., 5] = rng.integers(0, 4, size=(E, N))
        if tidy:
            side[..., 2] = np.where(side[..., 2] == -1, -2, side[..., 2])
            side[..., 3] = np.where(side[..., 3] == -1, -3, side[..., 3])
        side[~live] = -1
        # some dirty rows: qty <= 0 but not wiped, partial -1 fields
        dirty = rng.random((E, N)) < weird * 0.3
        side[..., 1] = np.where(dirty & live, rng.integers(-2, 1, size=(E, N)), side[..., 1])
        part = rng.random((E, N)) < weird * 0.3
        col = rng.integers(2, 6, size=(E, N))
        for c in range(2, 6):
            side[..., c] = np.where(part & live & (col == c), -1, side[..., c])
    msgs = np.zeros((E, M, 8), dtype=np.int32)
    msgs[..., 0] = rng.choice([1, 1, 1, 2, 2, 3, 4, 0, 5], size=(E, M))
    msgs[..., 1] = rng.choice([-1, 1, -1, 1, 0], size=(E, M))
    msgs[..., 2] = rng.integers(-3, 40, size=(E, M))
    msgs[..., 3] = rng.integers(97, 113, size=(E, M))
    msgs[..., 4] = rng.integers(-6, 12, size=(E, M))
    msgs[..., 5] = rng.integers(-4, 4, size=(E, M))
    msgs[..., 6] = rng.integers(0, 4, size=(E, M))
    msgs[..., 7] = rng.integers(0, 4, size=(E, M))
    w = rng.random((E, M, 8)) < weird * 0.2
    msgs = np.where(w, -1, msgs).astype(np.int32)
    z = rng.random((E, M)) < 0.05
    msgs[z] = 0                                   # all-zero rows are no-ops (JOBA:653)
    trades = np.full((E, T, 8), -1, dtype=np.int32)
    if tidy:
        msgs[..., 4] = np.where(msgs[..., 4] == -1, -2, msgs[..., 4])
        msgs[..., 5] = np.where(msgs[..., 5] == -1, -3, msgs[..., 5])
        msgs[z] = 0
        pre = np.arange(T)[None, :] < rng.integers(0, T + 1, size=(E, 1))
        trades[pre] = rng.integers(-1, 5, size=(int(pre.sum()), 8))
        trades[..., 4] = np.where(pre, np.abs(trades[..., 4]), -1)
    else:
        pre = rng.random((E, T)) < 0.3
        trades[pre] = rng.integers(-1, 5, size=(int(pre.sum()), 8))
    return asks, bids, trades, msgs


def synthetic_case(E, M, steps=1, seed=1234, N=100):
    """Initial books through the A8 recipe (oracle) + `steps` blocks of synthetic messages."""
    from oracle import c_oracle as C
    from vitmarl_b200 import synth
    l2 = synth.make_l2_books(E, seed)
    init = synth.init_msgs_from_l2_batched(l2)
    empty = np.full((E, N, 6), -1, dtype=np.int32)
    asks, bids, _, _, _ = C.lob_step(empty, empty.copy(), init, want_best=False)
    stream = synth.MessageStream(E, seed)
    return asks, bids, [stream.next(M) for _ in range(steps)]
