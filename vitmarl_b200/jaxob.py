"""Order-book ops with the reference's names and argument meaning, batched over a leading
environment axis (the form the reference reaches through ``jax.vmap(env.step)``), running
the hand-written sm_100a kernels through the C ABI.

Mirrors ``gymnax_exchange/jaxob/JaxOrderBookArrays.py`` (JOBA):
  scan_through_entire_array_save_bidask :720-752   scan_through_entire_array :665-685
  get_best_bid_and_ask_inclQuants :881-898          get_L2_state :1075-1106
  get_vision_L2_state :1108-1140                    init_orderside :901-911
  init_msgs_from_l2 :913-942

Arrays are ``torch.int32`` CUDA tensors (PyTorch is only the device-memory / stream
plumbing; a jax.ffi binding passes XLA buffers to the same C entry points).  ``key`` is
accepted and ignored exactly as the reference ignores it under the default cancel mode
(JOBA:130-137); cancel modes 2/3 raise VitmarlError(EUNSUPPORTED)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _capi
from .config import JAXLOB_Configuration

__all__ = ["init_orderside", "init_msgs_from_l2", "scan_through_entire_array",
           "scan_through_entire_array_save_bidask", "get_best_bid_and_ask_inclQuants",
           "get_L2_state", "get_vision_L2_state", "getCancelMsgs", "get_agent_trades", "agent_trade_stats", "filter_messages"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str, last: int) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise _capi.VitmarlError(_capi.ENODEVICE, f"{name} must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != torch.int32 or t.dim() != 3 or t.shape[-1] != last:
        raise _capi.VitmarlError(_capi.EINVAL, f"{name}: expected int32 [E,*,{last}], got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def init_orderside(nOrders: int = 100, E: int = 1, device="cuda") -> torch.Tensor:
    """JOBA:901-911, batched: int32 [E, nOrders, 6] filled with -1."""
    return torch.full((E, nOrders, 6), -1, dtype=torch.int32, device=device)


def init_msgs_from_l2(cfg: JAXLOB_Configuration, book_l2: torch.Tensor, time=None) -> torch.Tensor:
    """JOBA:913-942 batched: book_l2 int32 [E, 4L] -> limit messages int32 [E, 2L, 8]."""
    E, L4 = book_l2.shape
    L = L4 // 4
    data = book_l2.reshape(E, 2 * L, 2).to(torch.int32)
    out = torch.zeros((E, 2 * L, 8), dtype=torch.int32, device=book_l2.device)
    out[:, :, 3] = data[:, :, 0]
    out[:, :, 2] = data[:, :, 1]
    out[:, :, 0] = 1
    out[:, 0::2, 1] = -1
    out[:, 1::2, 1] = 1
    out[:, :, 4] = cfg.init_id
    out[:, :, 5] = cfg.init_id - torch.arange(2 * L, dtype=torch.int32, device=book_l2.device)[None, :]
    if time is None:
        time = (34200, 0)
    t = torch.as_tensor(time, dtype=torch.int32, device=book_l2.device).reshape(-1, 2)
    out[:, :, 6] = t[:, 0:1]
    out[:, :, 7] = t[:, 1:2]
    return out


def scan_through_entire_array_save_bidask(cfg: JAXLOB_Configuration, key, msg_array: torch.Tensor,
                                          book_state: Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]],
                                          N_steps: int, *, inplace: bool = False):
    """JOBA:720-752.  msg_array [E,M,8]; book_state = (asks [E,N,6], bids [E,N,6], trades [E,T,8] | None);
    -> ((asks, bids, trades), (best_asks [E,min(N_steps,M),2], best_bids [...])).
    ``trades=None`` means "all -1" (the only form used by marl_env.py:377) and skips one read."""
    del key
    asks, bids, trades = book_state
    asks, bids, msgs = _chk(asks, "asks", 6), _chk(bids, "bids", 6), _chk(msg_array, "msg_array", 8)
    E, N, _ = asks.shape
    M = msgs.shape[1]
    if trades is not None:
        trades = _chk(trades, "trades", 8)
        T = trades.shape[1]
    else:
        T = cfg.nTrades
    n_keep = max(0, min(int(N_steps), M))
    a_out = asks if inplace else torch.empty_like(asks)
    b_out = bids if inplace else torch.empty_like(bids)
    t_out = torch.empty((E, T, 8), dtype=torch.int32, device=asks.device)
    ba = torch.empty((E, n_keep, 2), dtype=torch.int32, device=asks.device)
    bb = torch.empty((E, n_keep, 2), dtype=torch.int32, device=asks.device)
    rc = _capi.lib().vitmarl_lob_step(_stream(), E, N, T, M, n_keep, _ptr(asks), _ptr(bids), _ptr(trades), _ptr(msgs),
                                      _ptr(a_out), _ptr(b_out), _ptr(t_out), _ptr(ba), _ptr(bb),
                                      int(cfg.cancel_mode), int(cfg.init_id))
    _capi.check(rc)
    return (a_out, b_out, t_out), (ba, bb)


def scan_through_entire_array(cfg: JAXLOB_Configuration, key, msg_array, book_state):
    """JOBA:665-685 -> (asks, bids, trades)"""
    del key
    asks, bids, trades = book_state
    asks, bids, msgs = _chk(asks, "asks", 6), _chk(bids, "bids", 6), _chk(msg_array, "msg_array", 8)
    E, N, _ = asks.shape
    M = msgs.shape[1]
    if trades is not None:
        trades = _chk(trades, "trades", 8)
        T = trades.shape[1]
    else:
        T = cfg.nTrades
    a_out, b_out = torch.empty_like(asks), torch.empty_like(bids)
    t_out = torch.empty((E, T, 8), dtype=torch.int32, device=asks.device)
    rc = _capi.lib().vitmarl_lob_step(_stream(), E, N, T, M, 0, _ptr(asks), _ptr(bids), _ptr(trades), _ptr(msgs),
                                      _ptr(a_out), _ptr(b_out), _ptr(t_out), None, None,
                                      int(cfg.cancel_mode), int(cfg.init_id))
    _capi.check(rc)
    return a_out, b_out, t_out


def get_best_bid_and_ask_inclQuants(cfg: JAXLOB_Configuration, askside, bidside):
    """JOBA:881-898 -> (best_ask [E,2], best_bid [E,2]) = [price, volume at price]"""
    del cfg
    asks, bids = _chk(askside, "askside", 6), _chk(bidside, "bidside", 6)
    E, N, _ = asks.shape
    ba = torch.empty((E, 2), dtype=torch.int32, device=asks.device)
    bb = torch.empty((E, 2), dtype=torch.int32, device=asks.device)
    _capi.check(_capi.lib().vitmarl_lob_best_bid_ask(_stream(), E, N, _ptr(asks), _ptr(bids), _ptr(ba), _ptr(bb)))
    return ba, bb


def get_L2_state(asks, bids, n_levels: int, cfg: JAXLOB_Configuration) -> torch.Tensor:
    """JOBA:1075-1106 -> int32 [E, 4*n_levels]"""
    asks, bids = _chk(asks, "asks", 6), _chk(bids, "bids", 6)
    E, N, _ = asks.shape
    l2 = torch.empty((E, 4 * n_levels), dtype=torch.int32, device=asks.device)
    _capi.check(_capi.lib().vitmarl_lob_render(_stream(), E, N, n_levels, 1, _ptr(asks), _ptr(bids), None,
                                               None, _ptr(l2), None, None, _capi.IMG_NONE, 0, 0))
    return l2


def get_vision_L2_state(asks, bids, n_levels: int, cfg: JAXLOB_Configuration) -> torch.Tensor:
    """JOBA:1108-1140 -> int32 [E, n_levels, 2 (price, vol), 2 (ask, bid)]"""
    asks, bids = _chk(asks, "asks", 6), _chk(bids, "bids", 6)
    E, N, _ = asks.shape
    raw = torch.empty((E, n_levels, 2, 2), dtype=torch.int32, device=asks.device)
    _capi.check(_capi.lib().vitmarl_lob_render(_stream(), E, N, n_levels, 1, _ptr(asks), _ptr(bids), None,
                                               _ptr(raw), None, None, None, _capi.IMG_NONE, 0, 0))
    return raw


def getCancelMsgs(bookside, agentID: int, size: int, side: int, cancel_time: torch.Tensor) -> torch.Tensor:
    """JOBA:756-782 batched: bookside [E,N,6], cancel_time int32 [E,2] (time_s, time_ns) -> int32 [E,size,8]."""
    book = _chk(bookside, "bookside", 6)
    E, N, _ = book.shape
    ct = cancel_time.to(torch.int32).contiguous()
    out = torch.empty((E, size, 8), dtype=torch.int32, device=book.device)
    _capi.check(_capi.lib().vitmarl_get_cancel_msgs(_stream(), E, N, size, _ptr(book), int(agentID), int(side), _ptr(ct), _ptr(out)))
    return out


def filter_messages(action_msgs, cnl_msgs):
    """`_filter_messages` batched (vision_env.py:622-684, mm_env.py:509-571): [E,n,8] x [E,n,8] -> (action_msgs, cnl_msgs)."""
    am = _chk(action_msgs, "action_msgs", 8)
    cm = _chk(cnl_msgs, "cnl_msgs", 8)
    if am.shape != cm.shape:
        raise ValueError("filter_messages: action and cancel messages must have the same shape")
    E, n, _ = am.shape
    ao, co = torch.empty_like(am), torch.empty_like(cm)
    _capi.check(_capi.lib().vitmarl_filter_messages(_stream(), E, n, _ptr(am), _ptr(cm), _ptr(ao), _ptr(co)))
    return ao, co


def agent_trade_stats(trades, agent_id: int, tick_size: int) -> torch.Tensor:
    """The trade reductions of the reward functions in one pass (vision_env.py:2076-2078, 2156-2163, 2191;
    mm_env.py:1906-1936): trades [E,T,8] -> int32 [E,8] = [sum qty, sum |qty|, c_rl, buyQuant, sellQuant, TradedVolume,
    inventory_delta, sum |qty| of the other executed trades]."""
    tr = _chk(trades, "trades", 8)
    E, T, _ = tr.shape
    out = torch.empty((E, 8), dtype=torch.int32, device=tr.device)
    _capi.check(_capi.lib().vitmarl_agent_trade_stats(_stream(), E, T, _ptr(tr), int(agent_id), int(tick_size), _ptr(out)))
    return out


def get_agent_trades(trades, agent_id: int) -> torch.Tensor:
    """JOBA:824-831 batched: trades [E,T,8] -> same shape, rows not involving agent_id (or not executed) zeroed."""
    tr = _chk(trades, "trades", 8)
    E, T, _ = tr.shape
    out = torch.empty_like(tr)
    _capi.check(_capi.lib().vitmarl_get_agent_trades(_stream(), E, T, _ptr(tr), int(agent_id), _ptr(out)))
    return out
