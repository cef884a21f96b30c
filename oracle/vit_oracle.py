"""ORACLE (test infrastructure, NOT product code) for the ViT encoder: a plain fp32 PyTorch
implementation of docs/VIT_SPEC.md operating on the flax-style parameter pytree.

PARITY UNPINNED BY THE REFERENCE: hiepday3324/ViT-MARL contains no ViT (SURVEY.md F2:
networks/vision_agent.py is a broken CNN stub, gate_fusion.py is empty), so there is no
reference implementation, golden vector or test to pin this against.  The architecture is the
standard pre-LN ViT encoder as flax would build it (nn.LayerNorm eps 1e-6, nn.gelu tanh
approximation, nn.MultiHeadDotProductAttention parameter layout); gradients come from autograd.
Tolerances used by the tests (bf16 tensor-core path vs this fp32 oracle):
  activations max|err| / max|ref| <= 2e-2;  gradients cosine >= 0.999 and rel-L2 <= 3e-2."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def vit_forward(cfg, params, x: torch.Tensor) -> torch.Tensor:
    """x [B,H,W,C] float32 -> [B,D] float32"""
    B = x.shape[0]
    P, D, h = cfg.patch, cfg.dim, cfg.heads
    dh = D // h
    Hp, Wp = cfg.img_h // P, cfg.img_w // P
    # flax nn.Conv(D, (P,P), strides=(P,P), padding='VALID') on NHWC == matmul of flattened (ph,pw,c) patches
    patches = x.reshape(B, Hp, P, Wp, P, cfg.channels).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp * Wp, P * P * cfg.channels)
    t = patches @ params["patch_embed"]["kernel"].reshape(P * P * cfg.channels, D) + params["patch_embed"]["bias"]
    t = t + params["pos_embed"]
    for l in range(cfg.depth):
        b = params[f"encoderblock_{l}"]
        a = b["MultiHeadDotProductAttention_0"]
        y = F.layer_norm(t, (D,), b["LayerNorm_0"]["scale"], b["LayerNorm_0"]["bias"], cfg.ln_eps)
        q = torch.einsum("btd,dhk->bhtk", y, a["query"]["kernel"]) + a["query"]["bias"][None, :, None, :]
        k = torch.einsum("btd,dhk->bhtk", y, a["key"]["kernel"]) + a["key"]["bias"][None, :, None, :]
        v = torch.einsum("btd,dhk->bhtk", y, a["value"]["kernel"]) + a["value"]["bias"][None, :, None, :]
        p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
        o = torch.einsum("bhtk,hkd->btd", p @ v, a["out"]["kernel"]) + a["out"]["bias"]
        t = t + o
        y = F.layer_norm(t, (D,), b["LayerNorm_1"]["scale"], b["LayerNorm_1"]["bias"], cfg.ln_eps)
        m = b["MlpBlock_0"]
        y = F.gelu(y @ m["Dense_0"]["kernel"] + m["Dense_0"]["bias"], approximate="tanh")
        t = t + y @ m["Dense_1"]["kernel"] + m["Dense_1"]["bias"]
    t = F.layer_norm(t, (D,), params["encoder_norm"]["scale"], params["encoder_norm"]["bias"], cfg.ln_eps)
    return t.mean(dim=1)


def tree_map(fn, tree):
    return {k: tree_map(fn, v) if isinstance(v, dict) else fn(v) for k, v in tree.items()}


def tree_leaves(tree, prefix=""):
    out = []
    for k, v in tree.items():
        out += tree_leaves(v, prefix + k + "/") if isinstance(v, dict) else [(prefix + k, v)]
    return out


def vit_value_and_grad(cfg, params, x, dy, want_dx=False):
    """-> y, grads pytree of <vit(x), dy> (and dx)"""
    p = tree_map(lambda t: t.detach().clone().float().requires_grad_(True), params)
    x = x.detach().clone().float().requires_grad_(want_dx)
    y = vit_forward(cfg, p, x)
    (y * dy.float()).sum().backward()
    g = tree_map(lambda t: t.grad, p)
    return y.detach(), g, (x.grad if want_dx else None)
