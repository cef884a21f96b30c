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

__all__ = ["BookState", "StepOutput", "reset", "step", "build_step_msgs", "auto_reset"]


@dataclasses.dataclass
class BookState:
    """The leaves of WorldState the hot path touches (StatesandParams.py:58-79)."""
    ask_raw_orders: torch.Tensor      # int32 [E,N,6]
    bid_raw_orders: torch.Tensor      # int32 [E,N,6]
    trades: torch.Tensor              # int32 [E,T,8]
    best_asks: torch.Tensor           # int32 [E,M,2]
    best_bids: torch.Tensor           # int32 [E,M,2]
    mid_price: torch.Tensor           # float32 [E]


@dataclasses.dataclass
class StepOutput:
    vision_obs: Optional[torch.Tensor]     # float32 [E,n,3,2]
    vision_raw: Optional[torch.Tensor]     # int32   [E,n,2,2]
    image: Optional[torch.Tensor]          # [E,H,W,2] bf16/u8


def reset(cfg: World_EnvironmentConfig, asks: torch.Tensor, bids: torch.Tensor, num_msgs_per_step: int) -> BookState:
    """World part of MARLEnv.reset_env (marl_env.py:186-190): tile the initial best bid/ask,
    mid = float32((best_bid + best_ask) / 2)."""
    ba, bb = get_best_bid_and_ask_inclQuants(cfg, asks, bids)
    E = asks.shape[0]
    mid = ((bb[:, 0] + ba[:, 0]).to(torch.float32) / 2.0)
    trades = torch.full((E, cfg.nTradesLogged, 8), -1, dtype=torch.int32, device=asks.device)
    return BookState(asks, bids, trades, ba[:, None, :].repeat(1, num_msgs_per_step, 1).contiguous(),
                     bb[:, None, :].repeat(1, num_msgs_per_step, 1).contiguous(), mid)


def step(cfg: World_EnvironmentConfig, state: BookState, msgs: torch.Tensor, *, n_levels: int = 10,
         want_obs: bool = True, want_raw: bool = False, image_hw=None, image_dtype=torch.bfloat16,
         inplace: bool = True, image_patch: Optional[int] = None) -> tuple[BookState, StepOutput]:
    """One fused env step for all E environments (msgs int32 [E,M,8]).
    ``image_patch=p`` (bf16 only) makes the kernel write the raster directly as the ViT's patch matrix
    ``[E, (H/p)*(W/p), p*p*2]`` -- the same values as the ``[E,H,W,2]`` image, in the order the patch-embedding GEMM reads
    them (``ViTEncoder.apply_packed(..., patches=True)``), so the encoder's patchify pass disappears."""
    asks, bids = _chk(state.ask_raw_orders, "asks", 6), _chk(state.bid_raw_orders, "bids", 6)
    msgs = _chk(msgs, "msgs", 8)
    E, N, _ = asks.shape
    M, T = msgs.shape[1], cfg.nTradesLogged
    dev = asks.device
    last_a = state.best_asks[:, -1, 0].contiguous()
    last_b = state.best_bids[:, -1, 0].contiguous()
    a_out = asks if inplace else torch.empty_like(asks)
    b_out = bids if inplace else torch.empty_like(bids)
    t_out = state.trades if (inplace and state.trades.shape == (E, T, 8)) else torch.empty((E, T, 8), dtype=torch.int32, device=dev)
    ba = torch.empty((E, M, 2), dtype=torch.int32, device=dev)
    bb = torch.empty((E, M, 2), dtype=torch.int32, device=dev)
    mid = torch.empty((E,), dtype=torch.float32, device=dev)
    raw = torch.empty((E, n_levels, 2, 2), dtype=torch.int32, device=dev) if want_raw else None
    norm = torch.empty((E, n_levels, 3, 2), dtype=torch.float32, device=dev) if want_obs else None
    img, code, H, W = None, _capi.IMG_NONE, 0, 0
    if image_hw is not None:
        H, W = image_hw
        code = {torch.bfloat16: _capi.IMG_BF16, torch.uint8: _capi.IMG_U8}[image_dtype]
        img = torch.empty((E, H, W, 2), dtype=image_dtype, device=dev)
        if image_patch is not None:
            if image_dtype != torch.bfloat16 or image_patch % 4 or H % image_patch or W % image_patch:
                raise _capi.VitmarlError(_capi.EINVAL, "image_patch needs a bf16 raster and a patch size that is a multiple of 4 dividing H and W")
            code = (int(image_patch) << 8) | _capi.IMG_BF16_PATCHES
            img = img.view(E, (H // image_patch) * (W // image_patch), image_patch * image_patch * 2)
    rc = _capi.lib().vitmarl_env_step(_stream(), E, N, T, M, _ptr(asks), _ptr(bids), _ptr(msgs), _ptr(last_a), _ptr(last_b),
                                      _ptr(a_out), _ptr(b_out), _ptr(t_out), _ptr(ba), _ptr(bb), _ptr(mid),
                                      n_levels, cfg.tick_size, _ptr(raw), None, _ptr(norm), _ptr(img), code, H, W,
                                      int(cfg.cancel_mode), int(cfg.init_id))
    _capi.check(rc)
    return BookState(a_out, b_out, t_out, ba, bb, mid), StepOutput(norm, raw, img)


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
               init_best_asks: torch.Tensor, init_best_bids: torch.Tensor, init_trades: Optional[torch.Tensor] = None) -> BookState:
    """Auto-reset of ``MARLEnv.step`` (``marl_env.py:737-766``) for the world-state leaves in :class:`BookState`, in place:
    environments with ``done`` take the books / trades of their sampled data window (``base_env.py:215-231``), the tiled
    initial best bid / ask and the initial mid price (``marl_env.py:186-190``); the others are untouched.
    ``window_index`` is the caller's ``jax.random.randint`` draw."""
    asks, bids = _chk(state.ask_raw_orders, "asks", 6), _chk(state.bid_raw_orders, "bids", 6)
    E, N, _ = asks.shape
    T, M = state.trades.shape[1], state.best_asks.shape[1]
    i32 = lambda t: None if t is None else t.to(torch.int32).contiguous()
    d, w = i32(done), i32(window_index)
    ia, ib, it = _chk(init_asks, "init_asks", 6), _chk(init_bids, "init_bids", 6), i32(init_trades)
    iba, ibb = i32(init_best_asks), i32(init_best_bids)
    if ia.shape[1] != N or ib.shape != ia.shape or iba.shape != (ia.shape[0], 2) or ibb.shape != iba.shape:
        raise _capi.VitmarlError(_capi.EINVAL, "auto_reset: init state shapes")
    rc = _capi.lib().vitmarl_auto_reset(_stream(), E, N, T, M, ia.shape[0], _ptr(d), _ptr(w), _ptr(ia), _ptr(ib), _ptr(it), _ptr(iba),
                                        _ptr(ibb), _ptr(asks), _ptr(bids), _ptr(state.trades), _ptr(state.best_asks),
                                        _ptr(state.best_bids), _ptr(state.mid_price))
    _capi.check(rc)
    return state
