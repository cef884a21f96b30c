"""Fused environment step: the book part of ``MARLEnv.step_env``
(``gymnax_exchange/jaxen/marl_env.py:377-393, 466-467, 637-662``) as ONE kernel launch:
order-book scan -> forward-filled best prices -> mid price -> vision observation
(+ optional H x W raster feeding the ViT).  State stays resident in HBM between steps."""
from __future__ import annotations

import dataclasses
from typing import Optional

import torch

from . import _capi
from .config import World_EnvironmentConfig
from .jaxob import _chk, _ptr, _stream, get_best_bid_and_ask_inclQuants

__all__ = ["BookState", "StepOutput", "StepBuffers", "reset", "step", "build_step_msgs", "auto_reset"]


@dataclasses.dataclass
class BookState:
    """The leaves of WorldState the hot path touches (StatesandParams.py:58-79)."""
    ask_raw_orders: torch.Tensor      # int32 [E,N,6]
    bid_raw_orders: torch.Tensor      # int32 [E,N,6]
    trades: torch.Tensor              # int32 [E,T,8]
    best_asks: torch.Tensor           # int32 [E,M,2]
    best_bids: torch.Tensor           # int32 [E,M,2]
    mid_price: torch.Tensor           # float32 [E]
    time: Optional[torch.Tensor] = None          # int32 [E,2] world clock (seconds, ns); None = not tracked
    delta_time: Optional[torch.Tensor] = None    # float32 [E] (marl_env.py:468), valid after a step when `time` is tracked


@dataclasses.dataclass
class StepOutput:
    vision_obs: Optional[torch.Tensor]     # float32 [E,n,3,2]
    vision_raw: Optional[torch.Tensor]     # int32   [E,n,2,2]
    image: Optional[torch.Tensor]          # [E,H,W,2] bf16/u8
    trade_stats: Optional[torch.Tensor] = None   # int32 [E,n_agents,8] (vitmarl_agent_trade_stats rows) when stat_agent_ids was given


def reset(cfg: World_EnvironmentConfig, asks: torch.Tensor, bids: torch.Tensor, num_msgs_per_step: int) -> BookState:
    """World part of MARLEnv.reset_env (marl_env.py:186-190): tile the initial best bid/ask,
    mid = float32((best_bid + best_ask) / 2).  (Reset-time setup, off the hot path: torch is used for the tiling only.)"""
    ba, bb = get_best_bid_and_ask_inclQuants(cfg, asks, bids)
    E = asks.shape[0]
    mid = ((bb[:, 0] + ba[:, 0]).to(torch.float32) / 2.0)
    trades = torch.full((E, cfg.nTradesLogged, 8), -1, dtype=torch.int32, device=asks.device)
    return BookState(asks, bids, trades, ba[:, None, :].repeat(1, num_msgs_per_step, 1).contiguous(),
                     bb[:, None, :].repeat(1, num_msgs_per_step, 1).contiguous(), mid)


class StepBuffers:
    """Output buffers of :func:`step`, allocated once per (E, M, options) and reused by every step of a rollout loop
    (the per-step path issues exactly one kernel launch and no allocation)."""

    def __init__(self):
        self.key = None
        self.t = {}

    def get(self, key, make):
        if self.key != key:
            self.key, self.t = key, make()
        return self.t


def step(cfg: World_EnvironmentConfig, state: BookState, msgs: torch.Tensor, *, n_levels: int = 10,
         want_obs: bool = True, want_raw: bool = False, image_hw=None, image_dtype=torch.bfloat16,
         inplace: bool = True, image_patch: Optional[int] = None, stat_agent_ids=None, keep_trades: bool = True,
         buffers: Optional[StepBuffers] = None) -> tuple[BookState, StepOutput]:
    """One fused env step for all E environments (msgs int32 [E,M,8]).
    ``image_patch=p`` (bf16 only) makes the kernel write the raster directly as the ViT's patch matrix
    ``[E, (H/p)*(W/p), p*p*2]`` -- the same values as the ``[E,H,W,2]`` image, in the order the patch-embedding GEMM reads
    them (``ViTEncoder.apply_packed(..., patches=True)``), so the encoder's patchify pass disappears.
    ``stat_agent_ids`` (up to 4 trader ids): the reward functions' integer trade reductions for these agents are computed from
    the step's trade log inside the kernel (``StepOutput.trade_stats``); with ``keep_trades=False`` the ``[T,8]`` log is then
    not written to HBM at all (``state.trades`` keeps its previous contents).
    The previous step's last best prices are read by stride from ``state.best_asks / best_bids`` (no gather launch); with
    ``inplace=True`` those tracks -- and with ``buffers`` every other output -- are rewritten in place."""
    asks, bids = _chk(state.ask_raw_orders, "asks", 6), _chk(state.bid_raw_orders, "bids", 6)
    msgs = _chk(msgs, "msgs", 8)
    E, N, _ = asks.shape
    M, T = msgs.shape[1], cfg.nTradesLogged
    dev = asks.device
    pa, pb = state.best_asks, state.best_bids          # [E, Mprev, 2]: last price of env e at element (e*Mprev + Mprev-1)*2
    if not (pa.is_contiguous() and pb.is_contiguous() and pa.shape == pb.shape and pa.dtype == torch.int32):
        raise _capi.VitmarlError(_capi.EINVAL, "state.best_asks / best_bids must be contiguous int32 [E,M,2]")
    Mprev = pa.shape[1]
    a_out = asks if inplace else torch.empty_like(asks)
    b_out = bids if inplace else torch.empty_like(bids)
    if not keep_trades and not stat_agent_ids:
        raise _capi.VitmarlError(_capi.EINVAL, "keep_trades=False needs stat_agent_ids (the trade log is reduced on chip instead)")
    t_out = state.trades if (inplace and state.trades.shape == (E, T, 8)) else torch.empty((E, T, 8), dtype=torch.int32, device=dev)
    same_tracks = inplace and Mprev == M
    ids = list(stat_agent_ids or [])
    if len(ids) > 4:
        raise _capi.VitmarlError(_capi.EINVAL, "at most 4 stat_agent_ids per launch")
    img_shape, code, H, W = None, _capi.IMG_NONE, 0, 0
    if image_hw is not None:
        H, W = image_hw
        code = {torch.bfloat16: _capi.IMG_BF16, torch.uint8: _capi.IMG_U8}[image_dtype]
        img_shape = (E, H, W, 2)
        if image_patch is not None:
            if image_dtype != torch.bfloat16 or image_patch % 4 or H % image_patch or W % image_patch:
                raise _capi.VitmarlError(_capi.EINVAL, "image_patch needs a bf16 raster and a patch size that is a multiple of 4 dividing H and W")
            code = (int(image_patch) << 8) | _capi.IMG_BF16_PATCHES
            img_shape = (E, (H // image_patch) * (W // image_patch), image_patch * image_patch * 2)

    def make():
        t = {"mid": torch.empty((E,), dtype=torch.float32, device=dev)}
        if not same_tracks:
            t["ba"] = torch.empty((E, M, 2), dtype=torch.int32, device=dev)
            t["bb"] = torch.empty((E, M, 2), dtype=torch.int32, device=dev)
        if want_raw:
            t["raw"] = torch.empty((E, n_levels, 2, 2), dtype=torch.int32, device=dev)
        if want_obs:
            t["norm"] = torch.empty((E, n_levels, 3, 2), dtype=torch.float32, device=dev)
        if img_shape is not None:
            t["img"] = torch.empty(img_shape, dtype=image_dtype, device=dev)
        if ids:
            t["stats"] = torch.empty((E, len(ids), 8), dtype=torch.int32, device=dev)
        if state.time is not None:
            t["dt"] = torch.empty((E,), dtype=torch.float32, device=dev)
        return t

    key = (E, M, N, T, n_levels, want_obs, want_raw, img_shape, image_dtype, len(ids), same_tracks, state.time is not None, str(dev))
    t = buffers.get(key, make) if buffers is not None else make()
    ba, bb = (pa, pb) if same_tracks else (t["ba"], t["bb"])
    a = _capi.EnvStepArgs()
    a.E, a.N, a.T, a.M = E, N, T, M
    a.asks_in, a.bids_in, a.msgs = _ptr(asks), _ptr(bids), _ptr(msgs)
    off = (Mprev - 1) * 2 * 4
    a.last_ask_price, a.last_bid_price, a.last_price_stride = pa.data_ptr() + off, pb.data_ptr() + off, 2 * Mprev
    a.asks_out, a.bids_out, a.trades_out = _ptr(a_out), _ptr(b_out), (_ptr(t_out) if keep_trades else None)
    a.best_asks, a.best_bids, a.mid_price = _ptr(ba), _ptr(bb), _ptr(t["mid"])
    a.n_levels, a.tick_size = n_levels, cfg.tick_size
    a.raw, a.l2, a.norm = _ptr(t.get("raw")), None, _ptr(t.get("norm"))
    a.image, a.img_dtype, a.H, a.W = _ptr(t.get("img")), code, H, W
    a.cancel_mode, a.init_id = int(cfg.cancel_mode), int(cfg.init_id)
    a.n_stat_agents = len(ids)
    for i, v in enumerate(ids):
        a.stat_agent_ids[i] = int(v)
    a.trade_stats = _ptr(t.get("stats"))
    tm = None
    if state.time is not None:                      # world clock: time <- the last message's time, delta_time (marl_env.py:468,482)
        tm = state.time
        if not (tm.is_cuda and tm.is_contiguous() and tm.dtype == torch.int32 and tuple(tm.shape) == (E, 2)):
            raise _capi.VitmarlError(_capi.EINVAL, "state.time must be a contiguous int32 [E,2] CUDA tensor")
        if not inplace:
            tm = torch.empty_like(tm)
        a.time_in, a.time_out, a.delta_time = _ptr(state.time), _ptr(tm), _ptr(t["dt"])
    import ctypes
    _capi.check(_capi.lib().vitmarl_env_step2(_stream(), ctypes.byref(a)))
    return (BookState(a_out, b_out, t_out, ba, bb, t["mid"], tm, t.get("dt")),
            StepOutput(t.get("norm"), t.get("raw"), t.get("img"), t.get("stats")))


def build_step_msgs(message_data: torch.Tensor, start_index: torch.Tensor, step_counter: torch.Tensor, n_data: int,
                    cancel_msgs: torch.Tensor, action_msgs: torch.Tensor, order_id_counter: torch.Tensor,
                    perm: Optional[torch.Tensor] = None, end_time_s: Optional[torch.Tensor] = None):
    """Message assembly of ``MARLEnv.step_env`` (``marl_env.py:272-344``, ``base_env.py:341-371``) for all envs:
    -> (combined int32 [E, Mc+Ma+n_data, 8], new_order_id_counter int32 [E]).  ``perm`` [E,Ma] are the indices of the
    caller's ``jax.random.permutation`` (None = no shuffle); ``end_time_s`` only for fixed_time episodes."""
    md = message_data.to(torch.int32).contiguous()
    E = start_index.shape[0]
    cm, am = _chk(cancel_msgs, "cancel_msgs", 8), _chk(action_msgs, "action_msgs", 8)
    Mc, Ma = cm.shape[1], am.shape[1]
    i32 = lambda t: None if t is None else t.to(torch.int32).contiguous()
    si, sc, oc, pm, et = i32(start_index), i32(step_counter), i32(order_id_counter), i32(perm), i32(end_time_s)
    out = torch.empty((E, Mc + Ma + n_data, 8), dtype=torch.int32, device=md.device)
    newc = torch.empty((E,), dtype=torch.int32, device=md.device)
    rc = _capi.lib().vitmarl_build_step_msgs(_stream(), E, md.shape[0], n_data, Mc, Ma, _ptr(md), _ptr(si), _ptr(sc), _ptr(et),
                                             _ptr(cm), _ptr(am), _ptr(pm), _ptr(oc), _ptr(out), _ptr(newc))
    _capi.check(rc)
    return out, newc


def auto_reset(state: BookState, done: torch.Tensor, window_index: torch.Tensor, init_asks: torch.Tensor, init_bids: torch.Tensor,
               init_best_asks: torch.Tensor, init_best_bids: torch.Tensor, init_trades: Optional[torch.Tensor] = None,
               check_window: bool = True) -> BookState:
    """Auto-reset of ``MARLEnv.step`` (``marl_env.py:737-766``) for the world-state leaves in :class:`BookState`, in place:
    environments with ``done`` take the books / trades of their sampled data window (``base_env.py:215-231``), the tiled
    initial best bid / ask and the initial mid price (``marl_env.py:186-190``); the others are untouched.
    ``window_index`` is the caller's ``jax.random.randint`` draw; the kernel never reads the init tables out of bounds -- a done
    environment with an index outside ``[0, n_windows)`` is left untouched and (``check_window=True``, one 4-byte D2H read)
    reported as an error."""
    for name in ("ask_raw_orders", "bid_raw_orders", "trades", "best_asks", "best_bids", "mid_price"):
        t = getattr(state, name)
        if not (t.is_cuda and t.is_contiguous()):          # the reset is written IN PLACE: a silent .contiguous() copy would lose it
            raise _capi.VitmarlError(_capi.EINVAL, f"auto_reset: state.{name} must be a contiguous CUDA tensor")
    asks, bids = _chk(state.ask_raw_orders, "asks", 6), _chk(state.bid_raw_orders, "bids", 6)
    E, N, _ = asks.shape
    T, M = state.trades.shape[1], state.best_asks.shape[1]
    i32 = lambda t: None if t is None else t.to(torch.int32).contiguous()
    d, w = i32(done), i32(window_index)
    ia, ib, it = _chk(init_asks, "init_asks", 6), _chk(init_bids, "init_bids", 6), i32(init_trades)
    iba, ibb = i32(init_best_asks), i32(init_best_bids)
    if ia.shape[1] != N or ib.shape != ia.shape or iba.shape != (ia.shape[0], 2) or ibb.shape != iba.shape:
        raise _capi.VitmarlError(_capi.EINVAL, "auto_reset: init state shapes")
    bad = torch.zeros(1, dtype=torch.int32, device=asks.device) if check_window else None
    rc = _capi.lib().vitmarl_auto_reset(_stream(), E, N, T, M, ia.shape[0], _ptr(d), _ptr(w), _ptr(ia), _ptr(ib), _ptr(it), _ptr(iba),
                                        _ptr(ibb), _ptr(asks), _ptr(bids), _ptr(state.trades), _ptr(state.best_asks),
                                        _ptr(state.best_bids), _ptr(state.mid_price), _ptr(bad))
    _capi.check(rc)
    if check_window and int(bad.item()):
        raise _capi.VitmarlError(_capi.EINVAL, f"auto_reset: window_index outside [0, {ia.shape[0]}) for a done environment (left untouched)")
    return state
