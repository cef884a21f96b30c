"""Observation render (stage 2) with the reference's names.

Mirrors ``ExecutionAgent._get_obs_vision`` / ``normalize_vision_obs``
(``gymnax_exchange/jaxen/vision_env.py:2709-2721, 2804-2854``) batched over environments, plus the
builder-defined raster of docs/RENDER_SPEC.md (the reference renders no image)."""
from __future__ import annotations

import torch

from . import _capi
from .jaxob import _chk, _ptr, _stream

__all__ = ["get_obs_vision", "normalize_vision_obs", "render_image"]


def get_obs_vision(asks, bids, mid_price: torch.Tensor, n_levels: int = 10, tick_size: int = 100,
                   normalize: bool = True) -> torch.Tensor:
    """vision_env.py:2709-2721: raw int32 [E,n,2,2] if not normalize else float32 [E,n,3,2]."""
    asks, bids = _chk(asks, "asks", 6), _chk(bids, "bids", 6)
    E, N, _ = asks.shape
    raw = torch.empty((E, n_levels, 2, 2), dtype=torch.int32, device=asks.device)
    norm = mid = None
    if normalize:
        mid = mid_price.to(torch.float32).contiguous()
        norm = torch.empty((E, n_levels, 3, 2), dtype=torch.float32, device=asks.device)
    _capi.check(_capi.lib().vitmarl_lob_render(_stream(), E, N, n_levels, tick_size, _ptr(asks), _ptr(bids), _ptr(mid),
                                               _ptr(raw), None, _ptr(norm), None, _capi.IMG_NONE, 0, 0))
    return norm if normalize else raw


def normalize_vision_obs(asks, bids, mid_price, n_levels: int = 10, tick_size: int = 100) -> torch.Tensor:
    """vision_env.py:2804-2854 (takes the book rather than the raw tensor: the kernel fuses
    get_vision_L2_state and the normalisation so the raw tensor never round-trips HBM)."""
    return get_obs_vision(asks, bids, mid_price, n_levels, tick_size, True)


def render_image(asks, bids, H: int = 64, W: int = 64, tick_size: int = 100, dtype=torch.bfloat16) -> torch.Tensor:
    """docs/RENDER_SPEC.md: [E,H,W,2] image (channel 0 ask, 1 bid) in {0,1}; bf16 or uint8."""
    asks, bids = _chk(asks, "asks", 6), _chk(bids, "bids", 6)
    E, N, _ = asks.shape
    code = {torch.bfloat16: _capi.IMG_BF16, torch.uint8: _capi.IMG_U8}[dtype]
    img = torch.empty((E, H, W, 2), dtype=dtype, device=asks.device)
    _capi.check(_capi.lib().vitmarl_lob_render(_stream(), E, N, 1, tick_size, _ptr(asks), _ptr(bids), None,
                                               None, None, None, _ptr(img), code, H, W))
    return img
