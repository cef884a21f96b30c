"""ctypes binding of oracle/lob_oracle.c (CPU ORACLE -- test infrastructure only).

Batched NumPy-array API mirroring the C-ABI of the product library so the parity tests
can call both with the same arguments."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblob_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lob_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.lob_oracle_log1p_f32.restype = ctypes.c_float
        _lib.lob_oracle_log1p_f32.argtypes = [ctypes.c_float]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads() -> int:
    return int(lib().lob_oracle_num_threads())


def lob_step(asks, bids, msgs, trades_in=None, T=100, n_keep=None, init_id=-2, want_best=True, nthreads=0):
    """-> asks[E,N,6], bids[E,N,6], trades[E,T,8], best_asks[E,n_keep,2], best_bids[E,n_keep,2]"""
    asks, bids, msgs = _i32(asks), _i32(bids), _i32(msgs)
    E, N, _ = asks.shape
    M = msgs.shape[1]
    if trades_in is not None:
        trades_in = _i32(trades_in)
        T = trades_in.shape[1]
    n_keep = M if n_keep is None else min(n_keep, M)
    a, b = np.empty_like(asks), np.empty_like(bids)
    t = np.empty((E, T, 8), dtype=np.int32)
    ba = np.empty((E, n_keep, 2), dtype=np.int32) if want_best else None
    bb = np.empty((E, n_keep, 2), dtype=np.int32) if want_best else None
    rc = lib().lob_oracle_step(E, N, T, M, n_keep, init_id, _p(asks), _p(bids), _p(trades_in), _p(msgs),
                               _p(a), _p(b), _p(t), _p(ba), _p(bb), nthreads)
    assert rc == 0
    return a, b, t, ba, bb


def best_bid_ask(asks, bids):
    asks, bids = _i32(asks), _i32(bids)
    E, N, _ = asks.shape
    ba, bb = np.empty((E, 2), np.int32), np.empty((E, 2), np.int32)
    lib().lob_oracle_best(E, N, _p(asks), _p(bids), _p(ba), _p(bb))
    return ba, bb


def render(asks, bids, mid_price=None, n_levels=10, tick=100, H=0, W=0, nthreads=0):
    """-> raw[E,n,2,2] i32, norm[E,n,3,2] f32 | None, img[E,H,W,2] u8 | None"""
    asks, bids = _i32(asks), _i32(bids)
    E, N, _ = asks.shape
    raw = np.empty((E, n_levels, 2, 2), np.int32)
    norm = mid = None
    if mid_price is not None:
        mid = np.ascontiguousarray(mid_price, dtype=np.float32)
        norm = np.empty((E, n_levels, 3, 2), np.float32)
    img = np.empty((E, H, W, 2), np.uint8) if H > 0 else None
    lib().lob_oracle_render(E, N, n_levels, tick, _p(asks), _p(bids), _p(mid), _p(raw), _p(norm), _p(img), H, W, nthreads)
    return raw, norm, img


def ffill_mid(best_asks, best_bids, last_ask_price, last_bid_price):
    ba, bb = _i32(best_asks).copy(), _i32(best_bids).copy()
    E, M, _ = ba.shape
    mid = np.empty((E,), np.float32)
    lib().lob_oracle_ffill_mid(E, M, _p(ba), _p(bb), _p(_i32(last_ask_price)), _p(_i32(last_bid_price)), _p(mid))
    return ba, bb, mid


def log1p_f32(x: float) -> np.float32:
    return np.float32(lib().lob_oracle_log1p_f32(ctypes.c_float(float(x))))
