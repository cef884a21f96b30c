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
        self._packed = None              # table packed by the last apply() (what vjp() differentiates)
        self._ws = None
        self._ws_shape = None
        self._fold_key = None            # (workspace, batch, caller's params_version) of the folded parameters left in the workspace
        # per-call switches handed to the library with every launch (the library keeps no global state): set e.g.
        # ``enc.options.fused = 0`` for the unfused kernel sequence or ``enc.options.timing = timing.handle``
        self.options = _capi.VitOptions.defaults()

    # ---- flax-like surface -------------------------------------------------------------------------
    def init(self, seed: int = 0, x: Optional[torch.Tensor] = None, device="cuda") -> Dict:
        return {"params": init_params(self.cfg, seed, device)}

    def apply(self, variables: Dict, x: torch.Tensor, *, train: bool = False) -> torch.Tensor:
        """x: [B,H,W,C] (bf16, or any dtype castable to it).  train=True keeps the activations
        needed by :meth:`vjp` in the workspace.  The fp32 pytree is re-packed (cast to the bf16 / fp32 table of the C ABI)
        on EVERY call: an optimiser that updates the master tensors in place keeps the same dict object, so caching on
        identity would serve stale weights.  Loops that know when their parameters change use :meth:`apply_packed`."""
        self._packed = pack_params(self.cfg, variables["params"])
        return self.apply_packed(self._packed, x, train=train)

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
                     params_version: Optional[int] = None, patches: bool = False) -> torch.Tensor:
        """``patches=True``: ``x`` is already the patch matrix ``[B, T, P*P*C]`` bf16 (the fused env step
        renders it that way, ``env.step(..., image_patch=P)``) and the patchify pass is skipped.
        ``params_version`` (inference only) is a caller-owned generation counter of the CONTENTS of ``packed`` (bump it after
        every optimiser step / in-place update).  When it equals the version of the previous call -- and that call used the
        same workspace and batch size -- the folded parameters left in the workspace are reused (the rollout loop between
        two optimiser updates); ``None`` (default) always folds again.  Nothing is keyed on Python object identity."""
        if not x.is_cuda:
            raise _capi.VitmarlError(_capi.ENODEVICE, "ViT input must be a CUDA tensor (there is no CPU fallback)")
        c = self.cfg
        if patches:
            want = (c.tokens, c.patch * c.patch * c.channels)
            if x.dim() != 3 or tuple(x.shape[1:]) != want or x.dtype != torch.bfloat16:
                raise _capi.VitmarlError(_capi.EINVAL, f"patches=True takes a bf16 patch matrix [B,{want[0]},{want[1]}], got {tuple(x.shape)}")
        elif x.dim() != 4 or tuple(x.shape[1:]) != (c.img_h, c.img_w, c.channels):
            raise _capi.VitmarlError(_capi.EINVAL, f"expected x [B,{c.img_h},{c.img_w},{c.channels}], got {tuple(x.shape)}")
        x = x.to(torch.bfloat16).contiguous()
        B = x.shape[0]
        shape = self._shape(B)
        ws = self._workspace(shape, train, x.device)
        y = out if out is not None else torch.empty((B, c.dim), dtype=torch.float32, device=x.device)
        ptrs = (ctypes.c_void_p * len(packed))(*[t.data_ptr() for t in packed])
        key = (ws.data_ptr(), B, params_version)
        mode = 1 if train else (2 if (params_version is not None and self._fold_key == key) else 0)
        if patches:
            mode |= _capi.VIT_INPUT_PATCHES
        rc = _capi.lib().vitmarl_vit_fwd_ex(torch.cuda.current_stream().cuda_stream, ctypes.byref(shape), ptrs, x.data_ptr(),
                                            y.data_ptr(), ws.data_ptr(), ws.numel(), mode, ctypes.byref(self.options))
        _capi.check(rc)
        self._fold_key = None if (train or params_version is None) else key   # a training pass lays the workspace out differently
        self._ws_shape = (B, train)
        return y

    def vjp(self, variables: Dict, dy: torch.Tensor, *, want_dx: bool = False):
        """Gradients w.r.t. the flax pytree (and optionally the input) for the last
        ``apply(..., train=True)`` call: the pullback of ``jax.vjp(module.apply, ...)``."""
        if self._packed is None:
            raise _capi.VitmarlError(_capi.EINVAL, "vjp needs a preceding apply(..., train=True)")
        grads, dx = self.vjp_packed(self._packed, dy, want_dx=want_dx)      # the table the forward pass ran on
        g = unpack_grads(self.cfg, grads)
        return (g, dx) if want_dx else g

    def num_buckets(self) -> int:
        """Gradient buckets of the backward pass (depth + 2), in completion order: encoder_norm, block L-1 .. block 0,
        patch_embed + pos_embed."""
        return self.cfg.depth + 2

    def bucket_param_ranges(self):
        """[(first, last+1)] index ranges into the packed table for each bucket of :meth:`num_buckets`."""
        L = self.cfg.depth
        return [(3 + 12 * L, 5 + 12 * L)] + [(3 + 12 * l, 15 + 12 * l) for l in range(L - 1, -1, -1)] + [(0, 3)]

    def vjp_packed(self, packed, dy: torch.Tensor, *, want_dx: bool = False, grads=None, flat: Optional[torch.Tensor] = None,
                   bucket_events=None, accumulate: bool = False):
        """``grads``: fp32 tensors to write into (default: fresh ones).  ``flat``: the ONE buffer those tensors are views of
        (zeroed with a single memset).  ``bucket_events``: ``torch.cuda.Event`` per bucket (:meth:`num_buckets`), recorded on the
        current stream as soon as that bucket's gradients are final -- ``parallel.GradAllReducer`` hangs the per-bucket
        all-reduce behind them.  ``accumulate=True`` adds this call's gradients to the contents of ``grads`` (micro-batches)."""
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
        opt = _capi.VitOptions.defaults(fused=self.options.fused, gemm_2cta=self.options.gemm_2cta, pdl=self.options.pdl,
                                        attn_flags=self.options.attn_flags, timing=self.options.timing,
                                        debug_timeline=self.options.debug_timeline)
        opt.accumulate = 1 if accumulate else 0
        if flat is not None:
            opt.grads_flat, opt.grads_flat_bytes = flat.data_ptr(), flat.numel() * flat.element_size()
        evs = None
        if bucket_events is not None:
            if len(bucket_events) != self.num_buckets():
                raise _capi.VitmarlError(_capi.EINVAL, f"bucket_events: need {self.num_buckets()} events")
            evs = (ctypes.c_void_p * len(bucket_events))(*[e.cuda_event for e in bucket_events])
            opt.bucket_events = ctypes.cast(evs, ctypes.c_void_p)
        rc = _capi.lib().vitmarl_vit_bwd_ex(torch.cuda.current_stream().cuda_stream, ctypes.byref(shape), ptrs, self._ws.data_ptr(),
                                            self._ws.numel(), dy.data_ptr(), gptrs, dx.data_ptr() if want_dx else None, ctypes.byref(opt))
        _capi.check(rc)
        return grads, dx
