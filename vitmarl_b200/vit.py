"""ViT encoder with a flax-style interface on top of the sm_100a kernels.

The reference reserves the slot (``gymnax_exchange/networks/vision_agent.py:6-40`` is a CNN
stub with invalid kwargs, ``gate_fusion.py`` is empty, the trainers carry ``# FIXME: APPLY VISION``
at ``jaxrl/MARL/ippo_rnn_JAXMARL.py:75,117,292,416``) but ships no ViT, so the architecture is the
builder's specification in docs/VIT_SPEC.md.  What is kept from the reference is the calling
convention: ``module.init(seed, x) -> {'params': pytree}``, ``module.apply({'params': p}, x[B,H,W,C])
-> [B, D]`` (NHWC, batch first, ``vision_agent.py:17``) and a flax-named parameter pytree
(``Dense.kernel [in, out]``, ``MultiHeadDotProductAttention_0/{query,key,value}/kernel [D, h, Dh]``,
``out/kernel [h, Dh, D]``) so checkpoints stay orbax-compatible.

PyTorch tensors are used only as device buffers; every FLOP runs in libvitmarl_b200.so."""
from __future__ import annotations

import ctypes
import dataclasses
import math
from typing import Dict, Optional

import torch

from . import _capi

__all__ = ["ViTConfig", "VIT_TINY_8", "VIT_SMALL_16", "ViTEncoder", "init_params", "pack_params", "unpack_grads"]


@dataclasses.dataclass(frozen=True)
class ViTConfig:
    img_h: int = 64
    img_w: int = 64
    channels: int = 2
    patch: int = 8
    dim: int = 192
    depth: int = 12
    heads: int = 3
    mlp_dim: int = 768
    ln_eps: float = 1e-6

    @property
    def tokens(self) -> int:
        return (self.img_h // self.patch) * (self.img_w // self.patch)

    @property
    def patch_dim(self) -> int:
        return self.patch * self.patch * self.channels


VIT_TINY_8 = ViTConfig(64, 64, 2, 8, 192, 12, 3, 768)            # BASELINE configs[1]
VIT_SMALL_16 = ViTConfig(128, 128, 2, 16, 384, 12, 6, 1536)      # BASELINE configs[3]
VIT_PARITY = ViTConfig(64, 64, 2, 8, 192, 2, 3, 768)             # BASELINE configs[0]: "tiny ViT (patch 8, depth 2)"


def init_params(cfg: ViTConfig, seed: int = 0, device="cuda") -> Dict:
    """flax-default initialisers: lecun-normal kernels, zero biases, N(0, 0.02) position embedding,
    LayerNorm scale 1 / bias 0.  fp32 master copy (what an optax optimiser would hold)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    D, h, dh = cfg.dim, cfg.heads, cfg.dim // cfg.heads

    def lecun(shape, fan_in):
        std = math.sqrt(1.0 / fan_in) / 0.87962566103423978     # flax truncated-normal variance correction
        return (torch.nn.init.trunc_normal_(torch.empty(shape), 0.0, 1.0, -2.0, 2.0, generator=g) * std).to(device)

    def z(*shape):
        return torch.zeros(shape, device=device)

    p = {"patch_embed": {"kernel": lecun((cfg.patch, cfg.patch, cfg.channels, D), cfg.patch_dim), "bias": z(D)},
         "pos_embed": (torch.randn((1, cfg.tokens, D), generator=g) * 0.02).to(device)}
    for l in range(cfg.depth):
        p[f"encoderblock_{l}"] = {
            "LayerNorm_0": {"scale": torch.ones(D, device=device), "bias": z(D)},
            "MultiHeadDotProductAttention_0": {
                "query": {"kernel": lecun((D, h, dh), D), "bias": z(h, dh)},
                "key": {"kernel": lecun((D, h, dh), D), "bias": z(h, dh)},
                "value": {"kernel": lecun((D, h, dh), D), "bias": z(h, dh)},
                "out": {"kernel": lecun((h, dh, D), D), "bias": z(D)}},
            "LayerNorm_1": {"scale": torch.ones(D, device=device), "bias": z(D)},
            "MlpBlock_0": {"Dense_0": {"kernel": lecun((D, cfg.mlp_dim), D), "bias": z(cfg.mlp_dim)},
                           "Dense_1": {"kernel": lecun((cfg.mlp_dim, D), cfg.mlp_dim), "bias": z(D)}}}
    p["encoder_norm"] = {"scale": torch.ones(D, device=device), "bias": z(D)}
    return p


def pack_params(cfg: ViTConfig, params: Dict):
    """flax pytree (fp32) -> packed table of the C ABI: bf16 [out, in] matrices, fp32 vectors."""
    D = cfg.dim
    bf = lambda t: t.to(torch.bfloat16).contiguous()
    f32 = lambda t: t.to(torch.float32).contiguous()
    out = [bf(params["patch_embed"]["kernel"].reshape(cfg.patch_dim, D).t()), f32(params["patch_embed"]["bias"]),
           f32(params["pos_embed"].reshape(cfg.tokens, D))]
    for l in range(cfg.depth):
        b = params[f"encoderblock_{l}"]
        a = b["MultiHeadDotProductAttention_0"]
        qkv_w = torch.cat([a[k]["kernel"].reshape(D, D).t() for k in ("query", "key", "value")], dim=0)    # [3D, D]
        qkv_b = torch.cat([a[k]["bias"].reshape(D) for k in ("query", "key", "value")], dim=0)
        out += [f32(b["LayerNorm_0"]["scale"]), f32(b["LayerNorm_0"]["bias"]), bf(qkv_w), f32(qkv_b),
                bf(a["out"]["kernel"].reshape(D, D).t()), f32(a["out"]["bias"]),
                f32(b["LayerNorm_1"]["scale"]), f32(b["LayerNorm_1"]["bias"]),
                bf(b["MlpBlock_0"]["Dense_0"]["kernel"].t()), f32(b["MlpBlock_0"]["Dense_0"]["bias"]),
                bf(b["MlpBlock_0"]["Dense_1"]["kernel"].t()), f32(b["MlpBlock_0"]["Dense_1"]["bias"])]
    out += [f32(params["encoder_norm"]["scale"]), f32(params["encoder_norm"]["bias"])]
    return out


def unpack_grads(cfg: ViTConfig, grads) -> Dict:
    """packed fp32 gradient table -> flax-shaped pytree (the inverse of pack_params)."""
    D, h, dh = cfg.dim, cfg.heads, cfg.dim // cfg.heads
    it = iter(grads)
    p = {"patch_embed": {"kernel": next(it).t().reshape(cfg.patch, cfg.patch, cfg.channels, D), "bias": next(it)},
         "pos_embed": next(it).reshape(1, cfg.tokens, D)}
    for l in range(cfg.depth):
        ln1s, ln1b, qkv_w, qkv_b, out_w, out_b, ln2s, ln2b, w1, b1, w2, b2 = (next(it) for _ in range(12))
        att = {}
        for i, k in enumerate(("query", "key", "value")):
            att[k] = {"kernel": qkv_w[i * D:(i + 1) * D].t().reshape(D, h, dh), "bias": qkv_b[i * D:(i + 1) * D].reshape(h, dh)}
        att["out"] = {"kernel": out_w.t().reshape(h, dh, D), "bias": out_b}
        p[f"encoderblock_{l}"] = {"LayerNorm_0": {"scale": ln1s, "bias": ln1b}, "MultiHeadDotProductAttention_0": att,
                                  "LayerNorm_1": {"scale": ln2s, "bias": ln2b},
                                  "MlpBlock_0": {"Dense_0": {"kernel": w1.t(), "bias": b1}, "Dense_1": {"kernel": w2.t(), "bias": b2}}}
    p["encoder_norm"] = {"scale": next(it), "bias": next(it)}
    return p


class ViTEncoder:
    """``apply({'params': p}, x) -> [B, D]`` and ``vjp`` through the C ABI."""

    def __init__(self, cfg: ViTConfig = VIT_TINY_8):
        self.cfg = cfg
        self._packed_src = None
        self._packed = None
        self._ws = None
        self._ws_shape = None
        self._fold_key = None

    # ---- flax-like surface -------------------------------------------------------------------------
    def init(self, seed: int = 0, x: Optional[torch.Tensor] = None, device="cuda") -> Dict:
        return {"params": init_params(self.cfg, seed, device)}

    def apply(self, variables: Dict, x: torch.Tensor, *, train: bool = False) -> torch.Tensor:
        """x: [B,H,W,C] (bf16, or any dtype castable to it).  train=True keeps the activations
        needed by :meth:`vjp` in the workspace."""
        packed = self._get_packed(variables["params"])
        return self.apply_packed(packed, x, train=train)

    # ---- packed fast path (what the rollout / bench use: no per-call re-packing) -----------------------
    def _get_packed(self, params: Dict):
        if self._packed_src is not params:
            self._packed, self._packed_src = pack_params(self.cfg, params), params
        return self._packed

    def _shape(self, B: int) -> _capi.VitShape:
        c = self.cfg
        return _capi.VitShape(B, c.img_h, c.img_w, c.channels, c.patch, c.dim, c.depth, c.heads, c.mlp_dim, c.ln_eps)

    def _workspace(self, shape: _capi.VitShape, train: bool, device) -> torch.Tensor:
        need = _capi.lib().vitmarl_vit_workspace_bytes(ctypes.byref(shape), int(train))
        if need == 0 and shape.batch > 0:
            raise _capi.VitmarlError(_capi.EINVAL, "unsupported ViT shape: " + str(self.cfg))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=device)
        return self._ws

    def apply_packed(self, packed, x: torch.Tensor, *, train: bool = False, out: Optional[torch.Tensor] = None,
                     params_unchanged: bool = False, patches: bool = False) -> torch.Tensor:
        """``patches=True`` (inference only): ``x`` is already the patch matrix ``[B, T, P*P*C]`` bf16 (the fused env step
        renders it that way, ``env.step(..., image_patch=P)``) and the patchify pass is skipped.
        ``params_unchanged=True`` (inference only) tells the library that the CONTENTS of ``packed`` are the same as in the
        previous call, so the folded parameters it left in the workspace are reused (the rollout loop between two optimiser
        updates).  It is honoured only when that previous call used the same workspace, batch size and table."""
        if not x.is_cuda:
            raise _capi.VitmarlError(_capi.ENODEVICE, "ViT input must be a CUDA tensor (there is no CPU fallback)")
        c = self.cfg
        if patches:
            want = (c.tokens, c.patch * c.patch * c.channels)
            if train or x.dim() != 3 or tuple(x.shape[1:]) != want or x.dtype != torch.bfloat16:
                raise _capi.VitmarlError(_capi.EINVAL, f"patches=True is for inference on a bf16 patch matrix [B,{want[0]},{want[1]}], got {tuple(x.shape)}")
        elif x.dim() != 4 or tuple(x.shape[1:]) != (c.img_h, c.img_w, c.channels):
            raise _capi.VitmarlError(_capi.EINVAL, f"expected x [B,{c.img_h},{c.img_w},{c.channels}], got {tuple(x.shape)}")
        x = x.to(torch.bfloat16).contiguous()
        B = x.shape[0]
        shape = self._shape(B)
        ws = self._workspace(shape, train, x.device)
        y = out if out is not None else torch.empty((B, c.dim), dtype=torch.float32, device=x.device)
        ptrs = (ctypes.c_void_p * len(packed))(*[t.data_ptr() for t in packed])
        key = (ws.data_ptr(), B, id(packed))
        mode = 1 if train else (2 if (params_unchanged and self._fold_key == key) else 0)
        if patches:
            mode |= _capi.VIT_INPUT_PATCHES
        rc = _capi.lib().vitmarl_vit_fwd(torch.cuda.current_stream().cuda_stream, ctypes.byref(shape), ptrs, x.data_ptr(),
                                         y.data_ptr(), ws.data_ptr(), ws.numel(), mode)
        _capi.check(rc)
        self._fold_key = None if train else key          # a training pass lays the workspace out differently
        self._ws_shape = (B, train)
        return y

    def vjp(self, variables: Dict, dy: torch.Tensor, *, want_dx: bool = False):
        """Gradients w.r.t. the flax pytree (and optionally the input) for the last
        ``apply(..., train=True)`` call: the pullback of ``jax.vjp(module.apply, ...)``."""
        packed = self._get_packed(variables["params"])
        grads, dx = self.vjp_packed(packed, dy, want_dx=want_dx)
        g = unpack_grads(self.cfg, grads)
        return (g, dx) if want_dx else g

    def vjp_packed(self, packed, dy: torch.Tensor, *, want_dx: bool = False, grads=None):
        if self._ws_shape is None or not self._ws_shape[1]:
            raise _capi.VitmarlError(_capi.EINVAL, "vjp needs a preceding apply(..., train=True)")
        B = self._ws_shape[0]
        c = self.cfg
        shape = self._shape(B)
        dy = dy.to(torch.float32).contiguous()
        if grads is None:
            grads = [torch.empty(t.shape, dtype=torch.float32, device=dy.device) for t in packed]
        dx = torch.empty((B, c.img_h, c.img_w, c.channels), dtype=torch.bfloat16, device=dy.device) if want_dx else None
        ptrs = (ctypes.c_void_p * len(packed))(*[t.data_ptr() for t in packed])
        gptrs = (ctypes.c_void_p * len(grads))(*[t.data_ptr() for t in grads])
        rc = _capi.lib().vitmarl_vit_bwd(torch.cuda.current_stream().cuda_stream, ctypes.byref(shape), ptrs, self._ws.data_ptr(),
                                         self._ws.numel(), dy.data_ptr(), gptrs, dx.data_ptr() if want_dx else None)
        _capi.check(rc)
        return grads, dx
