"""Default action -> message tables of the reference's two agents, for all E environments at once:
``ExecutionAgent._getActionMsgs_fixedQuant_complex`` (gymnax_exchange/jaxen/vision_env.py:1046-1142) and
``MarketMakingAgent._getActionMsgs_spread_skew`` (gymnax_exchange/jaxen/mm_env.py:1352-1491).  The other action spaces stay with the
caller.  ``best_asks`` / ``best_bids`` are the world state's best-price tracks ``[E, M, 2]``: the kernels read
``[-1][0]`` of every environment by stride, no gather."""
from __future__ import annotations

import torch

from . import _capi

__all__ = ["getActionMsgs_fixedQuant_complex", "getActionMsgs_spread_skew", "PLACEHOLDER_ORDER_ID"]

PLACEHOLDER_ORDER_ID = -9      # World_EnvironmentConfig.placeholder_order_id (jaxob_config.py:173)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _i32(t: torch.Tensor, name: str, shape) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous() and tuple(t.shape) == tuple(shape)):
        raise _capi.VitmarlError(_capi.EINVAL, f"{name} must be a contiguous int32 CUDA tensor of shape {tuple(shape)}")
    return t


def _last_price(track: torch.Tensor, name: str):
    if not (track.is_cuda and track.dtype == torch.int32 and track.is_contiguous() and track.dim() == 3 and track.shape[2] == 2):
        raise _capi.VitmarlError(_capi.EINVAL, f"{name} must be a contiguous int32 [E,M,2] CUDA tensor")
    M = track.shape[1]
    return track.data_ptr() + (M - 1) * 2 * 4, 2 * M


def getActionMsgs_fixedQuant_complex(action, best_asks, best_bids, is_sell_task, task_to_execute, quant_executed, time, trader_id: int,
                                     tick_size: int = 100, n_ticks_in_book: int = 1, fixed_quant_value: int = 10,
                                     time_delay_obs_act: int = 0, placeholder_order_id: int = PLACEHOLDER_ORDER_ID) -> torch.Tensor:
    """-> int32 [E, 4, 8]: limit orders at the far touch / mid / near touch / passive price with the action's quantities."""
    E = action.shape[0]
    pa, stride = _last_price(best_asks, "best_asks")
    pb, _ = _last_price(best_bids, "best_bids")
    out = torch.empty((E, 4, 8), dtype=torch.int32, device=action.device)
    rc = _capi.lib().vitmarl_exec_action_msgs_fixed_quants_complex(
        _stream(), E, _i32(action, "action", (E,)).data_ptr(), pa, pb, stride, _i32(is_sell_task, "is_sell_task", (E,)).data_ptr(),
        _i32(task_to_execute, "task_to_execute", (E,)).data_ptr(), _i32(quant_executed, "quant_executed", (E,)).data_ptr(),
        _i32(time, "time", (E, 2)).data_ptr(), int(trader_id), int(tick_size), int(n_ticks_in_book), int(fixed_quant_value),
        int(time_delay_obs_act), int(placeholder_order_id), out.data_ptr())
    _capi.check(rc)
    return out


def getActionMsgs_spread_skew(action, best_asks, best_bids, time, trader_id: int, tick_size: int = 100, spread_multiplier: float = 3.0,
                              skew_multiplier: float = 5.0, multiplier_type: str = "tick", fixed_quant_value: int = 10,
                              time_delay_obs_act: int = 0, placeholder_order_id: int = PLACEHOLDER_ORDER_ID) -> torch.Tensor:
    """-> int32 [E, 2, 8]: a bid and an ask limit order around the skewed mid price (actions 0..5 = spread type x skew type)."""
    E = action.shape[0]
    pa, stride = _last_price(best_asks, "best_asks")
    pb, _ = _last_price(best_bids, "best_bids")
    out = torch.empty((E, 2, 8), dtype=torch.int32, device=action.device)
    rc = _capi.lib().vitmarl_mm_action_msgs_spread_skew(
        _stream(), E, _i32(action, "action", (E,)).data_ptr(), pa, pb, stride, _i32(time, "time", (E, 2)).data_ptr(), int(trader_id),
        int(tick_size), float(spread_multiplier), float(skew_multiplier), {"tick": 0, "spread": 1}[multiplier_type],
        int(fixed_quant_value), int(time_delay_obs_act), int(placeholder_order_id), out.data_ptr())
    _capi.check(rc)
    return out
