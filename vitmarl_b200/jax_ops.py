"""jax.ffi wrappers mirroring the reference signatures over csrc/ffi/vitmarl_ffi.cc.

JAX is NOT installable in this image (no wheel, no network), so this module is written against the public ``jax.ffi`` API
(``register_ffi_target`` / ``ffi_call`` / ``custom_vjp``) and can only be exercised once JAX is present; importing it without
JAX raises ImportError with that explanation.  Everything the GPU tests verify goes through the identical C ABI via
``vitmarl_b200/_capi.py`` (ctypes), so the kernels behind these handlers are the tested ones; the handlers themselves add no
logic beyond unpacking buffers and attributes.

Drop-in use inside the reference (INTEGRATION.md has the full diff)::

    from vitmarl_b200 import jax_ops as vm
    # marl_env.py:377-384   (under jax.vmap(env.step): vmap_method="broadcast_all" turns the batch axis into E)
    (asks, bids, trades), (bestasks, bestbids) = vm.scan_through_entire_array_save_bidask(cfg, key, msgs, (asks, bids, trades), M)
    # ippo_rnn_JAXMARL.py:317 / :425
    enc = vm.vit_apply(vit_shape, packed_params, images)          # differentiable (custom_vjp -> vitmarl_vit_bwd)
"""
from __future__ import annotations

import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
except ImportError as e:          # pragma: no cover - this image has no JAX
    raise ImportError("vitmarl_b200.jax_ops needs JAX (not installable in this image: no wheel, no network); "
                      "use the ctypes binding vitmarl_b200._capi / the torch harness instead") from e

from . import _build

_FFI_SO = os.path.join(_build.LIB_DIR, "libvitmarl_ffi.so")
_registered = False


def _register():
    global _registered
    if _registered:
        return
    if not os.path.exists(_FFI_SO):
        _build.build_ffi(jax.ffi.include_dir())
    lib = ctypes.CDLL(_FFI_SO)
    for name in ("VitmarlLobStep", "VitmarlEnvStep", "VitmarlVitFwd", "VitmarlVitBwd"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def scan_through_entire_array_save_bidask(cfg, key, msg_array, book_state, N_steps):
    """job.scan_through_entire_array_save_bidask (JaxOrderBookArrays.py:720-752) -- `key` is accepted and ignored (the
    default cancel_mode consumes no randomness; modes 2/3 are rejected by the handler)."""
    _register()
    asks, bids, trades = book_state
    M = msg_array.shape[-2]
    out = (jax.ShapeDtypeStruct(asks.shape, jnp.int32), jax.ShapeDtypeStruct(bids.shape, jnp.int32),
           jax.ShapeDtypeStruct(trades.shape, jnp.int32),
           jax.ShapeDtypeStruct(asks.shape[:-2] + (N_steps, 2), jnp.int32), jax.ShapeDtypeStruct(asks.shape[:-2] + (N_steps, 2), jnp.int32))
    a, b, t, ba, bb = jax.ffi.ffi_call("VitmarlLobStep", out, vmap_method="broadcast_all")(
        msg_array, asks, bids, trades, n_keep=int(N_steps), cancel_mode=int(cfg.cancel_mode), init_id=int(cfg.init_id))
    return (a, b, t), (ba, bb)


def vit_apply(shape: dict, packed_params, x, mode: int = 1):
    """module.apply({'params': p}, x) with a VJP (ippo_rnn_JAXMARL.py:317, 423-475).  `shape` = VitmarlVitShape fields."""
    _register()
    from . import _capi
    s = _capi.VitShape(x.shape[0], shape["img_h"], shape["img_w"], shape["channels"], shape["patch"], shape["dim"], shape["depth"],
                       shape["heads"], shape["mlp_dim"], shape.get("ln_eps", 1e-6))
    ws_bytes = _capi.lib().vitmarl_vit_workspace_bytes(ctypes.byref(s), 1)
    attrs = dict(batch=x.shape[0], img_h=shape["img_h"], img_w=shape["img_w"], channels=shape["channels"], patch=shape["patch"],
                 dim=shape["dim"], depth=shape["depth"], heads=shape["heads"], mlp_dim=shape["mlp_dim"],
                 ln_eps=float(shape.get("ln_eps", 1e-6)))

    @jax.custom_vjp
    def f(params, x):
        y, _ = fwd(params, x)
        return y

    def fwd(params, x):
        y, ws = jax.ffi.ffi_call("VitmarlVitFwd", (jax.ShapeDtypeStruct((x.shape[0], shape["dim"]), jnp.float32),
                                                   jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8)))(x, *params, mode=mode, **attrs)
        return y, (params, ws)

    def bwd(res, dy):
        params, ws = res
        out = tuple(jax.ShapeDtypeStruct(p.shape, jnp.float32) for p in params)
        grads = jax.ffi.ffi_call("VitmarlVitBwd", out)(dy, ws, *params, **attrs)
        return tuple(grads), None

    f.defvjp(fwd, bwd)
    return f(tuple(packed_params), x)
